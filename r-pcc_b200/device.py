"""Stream-ordered batch stages over torch CUDA tensors (torch is plumbing: device memory and
streams; every kernel is in librpcc_b200.so).  One function per `rpcc_*_batch` entry point."""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.RpccError("expected CUDA tensors (no CPU fallback)")
        if t is not None and not t.is_contiguous():
            raise _lib.RpccError("expected contiguous tensors")


def project_batch(points, offsets, lidar, out=None):
    """points (N,3|4) f32 cuda, offsets (B+1,) int64 cuda -> range (B,H,W) f32."""
    _need_cuda(points, offsets)
    B = offsets.numel() - 1
    rng = out if out is not None else torch.empty((B, lidar.H, lidar.W), dtype=torch.float32, device=points.device)
    scratch = torch.empty((max(B, 1) * 4,), dtype=torch.int32, device=points.device)
    check(_lib.lib().rpcc_project_batch(ptr(points), points.shape[1], ptr(offsets), B, lidar.H, lidar.W,
                                        C.c_float(lidar.horizontal_FOV), C.c_float(lidar.vertical_max),
                                        C.c_float(lidar.vertical_min), ptr(rng), ptr(scratch), _stream()))
    return rng


def range_to_xyz_batch(rng, lut):
    _need_cuda(rng, lut)
    B = rng.shape[0]
    HW = lut.numel() // 3
    xyz = torch.empty((B,) + tuple(lut.shape), dtype=torch.float32, device=rng.device)
    check(_lib.lib().rpcc_range_to_xyz_batch(ptr(rng), ptr(lut), B, HW, ptr(xyz), _stream()))
    return xyz


def fps_batch(points, m):
    """points (B,n,3) f32 cuda -> idx (B,m) int32 (reference furthest_point_sample)."""
    _need_cuda(points)
    B, n, _ = points.shape
    temp = torch.empty((B, n), dtype=torch.float32, device=points.device)
    idx = torch.empty((B, m), dtype=torch.int32, device=points.device)
    check(_lib.lib().rpcc_fps_batch(ptr(points), B, n, m, ptr(temp), ptr(idx), _stream()))
    return idx


def segment_fps_batch(rng, lut, ground, m, thr=0.1):
    _need_cuda(rng, lut, ground)
    B, H, W = rng.shape[0], lut.shape[0], lut.shape[1]
    cidx = torch.empty((B, m), dtype=torch.int32, device=rng.device)
    centers = torch.empty((B, m, 3), dtype=torch.float32, device=rng.device)
    check(_lib.lib().rpcc_segment_fps_batch(ptr(rng), ptr(lut), ptr(ground), B, H, W, m, C.c_float(thr),
                                            ptr(cidx), ptr(centers), _stream()))
    return cidx, centers


def new_book(B, H, W, K, device):
    n = _lib.lib().rpcc_book_bytes(B, H, W, K)
    return torch.empty((n,), dtype=torch.uint8, device=device)


def assign_labels_batch(rng, lut, ground, centers, book=None):
    _need_cuda(rng, lut, ground, centers)
    B, H, W = rng.shape[0], lut.shape[0], lut.shape[1]
    m = centers.shape[1]
    labels = torch.empty((B, H, W), dtype=torch.uint8, device=rng.device)
    if book is None:
        book = new_book(B, H, W, m + 2, rng.device)
    check(_lib.lib().rpcc_assign_labels_batch(ptr(rng), ptr(lut), ptr(ground), ptr(centers), B, H, W, m,
                                              ptr(labels), ptr(book), _stream()))
    return labels, book


def label_stats_batch(rng, labels, K, book=None):
    _need_cuda(rng, labels)
    B, H, W = labels.shape
    if book is None:
        book = new_book(B, H, W, K, rng.device)
    check(_lib.lib().rpcc_label_stats_batch(ptr(rng), ptr(labels), B, H, W, K, ptr(book), _stream()))
    return book


def point_model_batch(rng, labels, ground, book, K):
    _need_cuda(rng, labels, ground, book)
    B, H, W = labels.shape
    model = torch.empty((B, K, 4), dtype=torch.float32, device=rng.device)
    results = torch.empty((B, 4), dtype=torch.int32, device=rng.device)
    check(_lib.lib().rpcc_point_model_batch(ptr(rng), ptr(labels), ptr(ground), ptr(book), B, H, W, K,
                                            ptr(model), ptr(results), _stream()))
    return model, results


def quantize_pack_batch(rng, labels, model, lut, book, step, step_per_label=None):
    _need_cuda(rng, labels, model, lut, book, step_per_label)
    B, H, W = labels.shape
    K = model.shape[1]
    HW = H * W
    symbols = torch.empty((B, HW), dtype=torch.int16, device=rng.device)
    contour = torch.empty((B, (HW + 7) // 8), dtype=torch.uint8, device=rng.device)
    seq = torch.empty((B, HW), dtype=torch.int16, device=rng.device)  # uint16 payload
    check(_lib.lib().rpcc_quantize_pack_batch(ptr(rng), ptr(labels), ptr(model), ptr(lut), ptr(book),
                                              ptr(step_per_label), C.c_float(step), B, H, W, K, ptr(symbols),
                                              C.c_size_t(HW), ptr(contour), ptr(seq), C.c_size_t(HW), None, None, _stream()))
    return symbols, contour, seq


def eval_batch(range_ref, range_rec, lut, lidar, labels_ref=None, labels_rec=None, f1_threshold=0.02, want_dist=False):
    """The --eval figures for B pairs of range images over one ray table (rpcc_eval_batch, csrc/evalq.cu).
    -> (metrics (B, 12) f64 cuda, dist1 (B,H,W) f32 | None, dist2 | None); dist* are the squared nearest-neighbour
    distances of the reference's chamfer kernel, -1 at pixels that are not points."""
    _need_cuda(range_ref, range_rec, lut, labels_ref, labels_rec)
    B = range_ref.shape[0]
    H, W = lidar.H, lidar.W
    from .batch import EVAL_COLS
    metrics = torch.empty((B, EVAL_COLS), dtype=torch.float64, device=range_ref.device)
    _lib.lib().rpcc_eval_workspace_bytes.restype = C.c_size_t
    ws = torch.empty((_lib.lib().rpcc_eval_workspace_bytes(B, H, W),), dtype=torch.uint8, device=range_ref.device)
    d1 = torch.empty((B, H, W), dtype=torch.float32, device=range_ref.device) if want_dist else None
    d2 = torch.empty((B, H, W), dtype=torch.float32, device=range_ref.device) if want_dist else None
    check(_lib.lib().rpcc_eval_batch(ptr(range_ref), ptr(range_rec), ptr(lut), ptr(labels_ref), ptr(labels_rec), B, H, W,
                                     C.c_double(lidar.horizontal_FOV), C.c_double(lidar.vertical_max),
                                     C.c_double(lidar.vertical_min), C.c_float(f1_threshold ** 2), ptr(d1), ptr(d2),
                                     ptr(metrics), ptr(ws), _stream()))
    return metrics, d1, d2
