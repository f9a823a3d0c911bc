"""Batched encoder / decoder: many frames per kernel launch, host entropy coding on threads.

This is the re-plumbed tools/compress_datalist.py (reference :48-206: a ThreadPoolExecutor over a
per-frame closure that holds the GIL in every native call): frames go to the GPU in batches
(rpcc_encoder_encode_host pipelines upload / kernels / download over stream slots), the five
`.rpcc` sections come back packed, and bz2/deflate -- sequential by nature -- run on a host
thread pool while the next batch is on the device.
"""
import concurrent.futures as futures
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr
from .compress_utils import BasicCompressor, pack_bitstream, parse_bitstream
from .config import load_compressor_cfg
from .lidar import LidarConfig


class _EncoderConfig(C.Structure):
    _fields_ = [("H", C.c_int), ("W", C.c_int), ("hfov", C.c_double), ("vmax", C.c_double), ("vmin", C.c_double),
                ("cluster_num", C.c_int), ("ground_threshold", C.c_float), ("step", C.c_double),
                ("nonuniform", C.c_int), ("level_num", C.c_int), ("level_kp_num", C.c_int * 8),
                ("level_dacc", C.c_double * 8), ("ground_level", C.c_int), ("feature_region", C.c_int),
                ("segments", C.c_int), ("sharp_num", C.c_int), ("less_sharp_num", C.c_int), ("flat_num", C.c_int),
                ("max_batch", C.c_int), ("max_points", C.c_int64), ("device", C.c_int),
                ("model_method", C.c_int), ("plane_angle_threshold", C.c_float), ("host_chunk", C.c_int)]


RESULT_DTYPE = np.dtype([("sym_count", np.uint32), ("seq_count", np.uint32), ("model_rows", np.uint32),
                         ("flags", np.uint32)])


def _pinned(shape, dtype):
    return torch.empty(shape, dtype=dtype, pin_memory=True)


def default_workers():
    """Host entropy-coder threads of one process: the box's cores split evenly over the ranks of a torchrun launch
    (LOCAL_WORLD_SIZE), one left for the thread that feeds the GPU."""
    n = os.cpu_count() or 2
    try:
        ranks = int(os.environ.get("LOCAL_WORLD_SIZE") or os.environ.get("WORLD_SIZE") or 1)
    except ValueError:
        ranks = 1
    return max(1, n // max(ranks, 1) - 1)


class BatchEncoder:
    """project -> ground fit -> FPS -> labels -> (key points) -> point (or plane) models -> quantise + pack for
    batches of frames.  `accuracy` is the yaml value (step = 2 * accuracy, tools/compress.py:46)."""

    def __init__(self, lidar="Velodyne64E", accuracy=None, nonuniform=None, compressor_cfg=None, max_batch=256,
                 max_points=None, device=None, basic_compressor=None, workers=None, model_method=None, host_chunk=0):
        cfg = load_compressor_cfg(compressor_cfg) if not isinstance(compressor_cfg, dict) else compressor_cfg
        self.cfg = cfg
        self.model_method = model_method or cfg["modeling_method"]
        self.lidar = lidar if isinstance(lidar, LidarConfig) else LidarConfig(lidar)
        if cfg["segment_method"] != "FPS":
            raise NotImplementedError("only FPS segmentation is on this path")
        if self.model_method not in ("point", "plane"):
            raise ValueError("model_method must be 'point' or 'plane'")
        self.accuracy = cfg["accuracy"] if accuracy is None else accuracy
        self.step = self.accuracy * 2
        self.uniform = (cfg["compress_framework"] == "uniform") if nonuniform is None else (not nonuniform)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.max_batch = int(max_batch)
        self.max_points = int(max_points) if max_points else self.max_batch * 140000
        self.K = int(cfg["cluster_num"]) + 2
        self.method = basic_compressor or cfg["basic_compressor"]
        self.workers = workers or default_workers()
        c = _EncoderConfig()
        c.H, c.W = self.lidar.H, self.lidar.W
        c.hfov, c.vmax, c.vmin = self.lidar.horizontal_FOV, self.lidar.vertical_max, self.lidar.vertical_min
        c.cluster_num = int(cfg["cluster_num"])
        c.ground_threshold = float(cfg["ground_threshold"])
        c.step = float(self.step)
        c.nonuniform = 0 if self.uniform else 1
        lk, ld = list(cfg["level_key_point_num"]), list(cfg["level_delta_acc"])
        c.level_num = len(lk)
        for i in range(len(lk)):
            c.level_kp_num[i] = int(lk[i])
            c.level_dacc[i] = float(ld[i])
        c.ground_level = int(cfg["ground_salience_level"])
        c.feature_region = int(cfg["feature_region"])
        c.segments = int(cfg["segments"])
        c.sharp_num = int(cfg["sharp_num"])
        c.less_sharp_num = int(cfg["less_sharp_num"])
        c.flat_num = int(cfg["flat_num"])
        c.max_batch = self.max_batch
        c.max_points = self.max_points
        c.device = self.device
        c.model_method = 1 if self.model_method == "plane" else 0
        c.plane_angle_threshold = float(cfg["plane_angle_threshold"])
        c.host_chunk = int(host_chunk)
        self._c = c
        self._h = C.c_void_p(0)
        check(_lib.lib().rpcc_encoder_create(C.byref(c), C.byref(self._h)))
        self._pool = None
        self._host = None

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if self._h:
            _lib.lib().rpcc_encoder_destroy(self._h)
            self._h = C.c_void_p(0)
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ device path
    @property
    def slots(self):
        return int(_lib.lib().rpcc_encoder_slots())

    def encode_device(self, slot, points, offsets, B, ground=None):
        """points (N,3|4) f32 cuda, offsets (B+1,) int64 cuda, ground (B,4) f32 cuda or None.  Asynchronous."""
        check(_lib.lib().rpcc_encoder_encode_device(self._h, int(slot), ptr(points), int(points.shape[1]), ptr(offsets),
                                                    int(B), ptr(ground) if ground is not None else None))

    def sync(self):
        check(_lib.lib().rpcc_encoder_sync(self._h))

    STAGES = ("project", "ground", "fps", "assign", "keypoints", "model", "quantize")

    def profile(self, enable=True):
        """Start (or stop) recording CUDA events between the stages of every chain call."""
        check(_lib.lib().rpcc_encoder_profile(self._h, 1 if enable else 0))

    def stage_times(self):
        """-> ({stage: summed ms}, frames covered, chain calls) since profile(True)."""
        ms = (C.c_double * 7)()
        frames = C.c_longlong(0)
        calls = C.c_int(0)
        check(_lib.lib().rpcc_encoder_stage_times(self._h, ms, C.byref(frames), C.byref(calls)))
        return dict(zip(self.STAGES, list(ms))), int(frames.value), int(calls.value)

    def stream(self, slot):
        _lib.lib().rpcc_encoder_stream.restype = C.c_void_p
        return _lib.lib().rpcc_encoder_stream(self._h, int(slot))

    def device_buffer(self, slot, name, shape, dtype):
        """A torch view of one of the slot's device buffers (tests / chaining)."""
        _lib.lib().rpcc_encoder_device_buffer.restype = C.c_void_p
        p = _lib.lib().rpcc_encoder_device_buffer(self._h, int(slot), name.encode())
        if not p:
            raise KeyError(name)
        n = int(np.prod(shape))
        itemsize = torch.empty((), dtype=dtype).element_size()
        # wrap the raw device pointer without copying
        iface = {"shape": (n * itemsize,), "typestr": "|u1", "data": (p, False), "version": 3}
        holder = type("_Dev", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device="cuda:%d" % self.device).view(dtype).view(*shape)

    # ------------------------------------------------------------------ host path
    def _host_buffers(self, B, npts_cap):
        HW, K = self.lidar.HW, self.K
        h = self._host
        if h is None or h["B"] < B:
            cap = max(B, self.max_batch)
            h = dict(B=cap,
                     results=_pinned((cap, 4), torch.int32),
                     model=_pinned((cap, K, 4), torch.float32),
                     contour=_pinned((cap, (HW + 7) // 8), torch.uint8),
                     seq=_pinned((cap * HW,), torch.int16),
                     symbols=_pinned((cap * HW,), torch.int16),
                     salience=_pinned((cap, K), torch.uint8))
            self._host = h
        return h

    def encode_host(self, points, offsets, grounds=None):
        """points: (N,3|4) f32 numpy or pinned torch tensor (all frames back to back); offsets (B+1,) int64;
        grounds (B,4) or None (fit on device).  Returns a dict of numpy views (valid until the next call):
        results (structured: sym_count, seq_count, model_rows, flags), model (B,K,4) f32, contour (B,HW/8) u8,
        seq (sum,) u16, symbols (sum,) i16, salience (B,K) u8 | None, plus sym_off / seq_off (B+1,)."""
        pts = points if isinstance(points, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32))
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        B = off.size - 1
        h = self._host_buffers(B, pts.shape[0])
        g = None
        if grounds is not None:
            g = np.ascontiguousarray(grounds, dtype=np.float32)
        HW = self.lidar.HW
        check(_lib.lib().rpcc_encoder_encode_host(
            self._h, ptr(pts), int(pts.shape[1]), ptr(off), B, ptr(g) if g is not None else None, ptr(h["results"]),
            ptr(h["model"]), ptr(h["contour"]), ptr(h["seq"]), C.c_size_t(h["B"] * HW), ptr(h["symbols"]),
            C.c_size_t(h["B"] * HW), ptr(h["salience"]) if not self.uniform else None))
        res = h["results"].numpy()[:B].view(np.uint32).copy().view(RESULT_DTYPE).reshape(B)
        sym_off = np.concatenate(([0], np.cumsum(res["sym_count"], dtype=np.int64)))
        seq_off = np.concatenate(([0], np.cumsum(res["seq_count"], dtype=np.int64)))
        return dict(results=res, model=h["model"].numpy()[:B], contour=h["contour"].numpy()[:B],
                    seq=h["seq"].numpy()[:seq_off[-1]].view(np.uint16), symbols=h["symbols"].numpy()[:sym_off[-1]],
                    salience=None if self.uniform else h["salience"].numpy()[:B], sym_off=sym_off, seq_off=seq_off)

    @staticmethod
    def frame_sections(enc, b):
        """The five uncompressed sections of frame b, as compress_point_cloud builds them
        (utils/compress_utils.py:141-161)."""
        r = enc["results"][b]
        rows = int(r["model_rows"])
        sec = {}
        if enc["salience"] is not None:
            sec["salience_level"] = enc["salience"][b, :rows].tobytes()
        sec["contour_map"] = enc["contour"][b].tobytes()
        sec["idx_sequence"] = enc["seq"][enc["seq_off"][b]:enc["seq_off"][b + 1]].tobytes()
        sec["plane_param"] = enc["model"][b, :rows].tobytes()
        sec["residual_quantized"] = enc["symbols"][enc["sym_off"][b]:enc["sym_off"][b + 1]].tobytes()
        return sec

    def _entropy(self, sections):
        bc = BasicCompressor(method_name=self.method, gzip_mtime=0)
        return pack_bitstream({k: bc.compress(v, section=k) for k, v in sections.items()}, uniform=self.uniform)

    def compress(self, points, offsets, grounds=None):
        """-> list of B `.rpcc` byte strings (GPU stages + host entropy coding on the thread pool)."""
        enc = self.encode_host(points, offsets, grounds)
        B = len(enc["results"])
        if self._pool is None:
            self._pool = futures.ThreadPoolExecutor(self.workers)
        secs = [self.frame_sections(enc, b) for b in range(B)]
        return list(self._pool.map(self._entropy, secs))


class BatchDecoder:
    """.rpcc byte strings -> reconstructed range images / point clouds, B frames per launch
    (tools/decompress.py:88-112 for many frames)."""

    def __init__(self, lidar="Velodyne64E", accuracy=None, nonuniform=None, compressor_cfg=None, basic_compressor=None,
                 workers=None):
        cfg = load_compressor_cfg(compressor_cfg) if not isinstance(compressor_cfg, dict) else compressor_cfg
        self.cfg = cfg
        self.lidar = lidar if isinstance(lidar, LidarConfig) else LidarConfig(lidar)
        self.accuracy = cfg["accuracy"] if accuracy is None else accuracy
        self.step = self.accuracy * 2
        self.uniform = (cfg["compress_framework"] == "uniform") if nonuniform is None else (not nonuniform)
        self.method = basic_compressor or cfg["basic_compressor"]
        self.level_acc = np.array([self.step] * len(cfg["level_key_point_num"])) + np.array(cfg["level_delta_acc"])
        self.workers = workers or default_workers()
        self._lut = None

    def _unpack(self, blob):
        bc = BasicCompressor(method_name=self.method)
        return bc.decompress_dict(parse_bitstream(blob, uniform=self.uniform))

    def decode(self, blobs, want_xyz=True):
        """-> dict(range (B,H,W) f32 cuda, xyz (B,H,W,3) f32 cuda | None, labels (B,H,W) u8 cuda, results)"""
        B = len(blobs)
        H, W, HW = self.lidar.H, self.lidar.W, self.lidar.HW
        with futures.ThreadPoolExecutor(self.workers) as pool:
            secs = list(pool.map(self._unpack, blobs))
        rows = [len(s["plane_param"]) // 16 for s in secs]
        K = max(max(rows), 2)
        cb = (HW + 7) // 8
        seq_n = np.array([len(s["idx_sequence"]) // 2 for s in secs], np.uint32)
        sym_n = np.array([len(s["residual_quantized"]) // 2 for s in secs], np.uint32)
        seq_stride, sym_stride = int(max(seq_n.max(), 1)), int(max(sym_n.max(), 1))
        contour = np.zeros((B, cb), np.uint8)
        seq = np.zeros((B, seq_stride), np.uint16)
        sym = np.zeros((B, sym_stride), np.int16)
        model = np.zeros((B, K, 4), np.float32)
        steps = np.full((B, K), self.step, np.float64)
        for b, s in enumerate(secs):
            c = np.frombuffer(s["contour_map"], np.uint8)
            contour[b, :min(cb, c.size)] = c[:cb]
            seq[b, :seq_n[b]] = np.frombuffer(s["idx_sequence"], np.uint16)
            sym[b, :sym_n[b]] = np.frombuffer(s["residual_quantized"], np.int16)
            model[b, :rows[b]] = np.frombuffer(s["plane_param"], np.float32).reshape(-1, 4)
            if not self.uniform:
                sal = np.frombuffer(s["salience_level"], np.uint8)
                steps[b, :sal.size] = self.level_acc[sal]
        dev = torch.device("cuda", torch.cuda.current_device())
        if self._lut is None:
            self._lut = torch.from_numpy(self.lidar.transform_map()).to(dev)
        t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
        d_contour, d_model, d_steps = t(contour), t(model), t(steps)
        d_seq, d_sym = t(seq.view(np.int16)), t(sym)
        d_seq_n, d_sym_n = t(seq_n.view(np.int32)), t(sym_n.view(np.int32))
        labels = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        rng = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        xyz = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev) if want_xyz else None
        book = torch.empty((_lib.lib().rpcc_book_bytes(B, H, W, K),), dtype=torch.uint8, device=dev)
        results = torch.empty((B, 4), dtype=torch.int32, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(_lib.lib().rpcc_decode_batch(ptr(d_contour), ptr(d_seq), C.c_size_t(seq_stride), ptr(d_seq_n), ptr(d_sym),
                                           C.c_size_t(sym_stride), ptr(d_sym_n), ptr(d_model), ptr(d_steps),
                                           ptr(self._lut), B, H, W, K, ptr(labels), ptr(rng),
                                           ptr(xyz) if want_xyz else None, ptr(book), ptr(results), st))
        res = results.cpu().numpy().view(np.uint32).copy().view(RESULT_DTYPE).reshape(B)
        bad = [b for b in range(B) if res["sym_count"][b] != sym_n[b] or res["seq_count"][b] != seq_n[b] or (res["flags"][b] & 6)]
        if bad:
            raise ValueError("malformed .rpcc stream(s) at batch index %s (section lengths do not match the label map; "
                             "wrong lidar / framework settings?)" % bad[:8])
        return dict(range=rng, xyz=xyz, labels=labels, results=res)
