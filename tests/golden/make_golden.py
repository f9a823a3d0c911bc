"""Generates tests/golden/example_golden.npz by running the REFERENCE's own Python + C++ on
assets/example_data/example.bin.  Runs only in the dev container (needs /root/reference and
oracle/_ref); the GPU box uses the committed .npz.

What comes from where:
  * dataset.load_range_image_points_from_file, PointCloudSegment.cluster_modeling / intra_predict,
    QuantizationModule.quantize_residual, compress_point_cloud, save_compressed_bitstream,
    read_compressed_bitstream, decompress_point_cloud, dequantize_residual: the reference's own
    utils/*.py and dataset/*.py, imported from /root/reference, calling its own C++
    (ops/cpp_modules/src/cpp_modules.cpp compiled into oracle/_ref).
  * PointCloudSegment.segment needs CUDA (torch ops + the FPS kernel) and open3d: the ground model is
    injected (tests/conftest.py EXAMPLE_GROUND) and seg_idx comes from the oracle's restatement of
    segment(); tests/test_gpu_stages.py pins that restatement on the GPU box against torch's own ops
    and the reference's compiled FPS kernel.
Large arrays are stored as sha256 digests, small ones in full.
The non-uniform run is executed with glibc's MALLOC_PERTURB_=255 (see __main__): the reference
reads uninitialised memory there and is otherwise not reproducible.
"""
import hashlib
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "stubs"), os.path.join(ROOT, "oracle", "_ref"), REF, ROOT]

import ops                                            # noqa: E402  (namespace package: oracle/_ref/ops + reference ops)
fps_pkg = types.ModuleType("ops.fps")                 # CUDA-only; segment() is not called here
fps_stub = types.ModuleType("ops.fps.fps_utils")
fps_pkg.fps_utils = fps_stub
ops.fps = fps_pkg
sys.modules["ops.fps"] = fps_pkg
sys.modules["ops.fps.fps_utils"] = fps_stub

import oracle                                                            # noqa: E402
from dataset import build_dataset                                       # noqa: E402  (reference)
from utils.compress_utils import (BasicCompressor, QuantizationModule, compress_point_cloud,  # noqa: E402
                                  decompress_point_cloud, read_compressed_bitstream, save_compressed_bitstream)
from utils.segment_utils import PointCloudSegment                       # noqa: E402
from utils.utils import load_compressor_cfg                             # noqa: E402

GROUND = np.array([0.0057, -0.0445, -0.9990, -1.7938])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run(nonuniform):
    cfg = load_compressor_cfg(os.path.join(REF, "cfgs/compressor.yaml"))
    accuracy = cfg["accuracy"] * 2
    bc = BasicCompressor(compressor_yaml=os.path.join(REF, "cfgs/compressor.yaml"))
    dataset = build_dataset(lidar_type="Velodyne64E")
    pc_seg = PointCloudSegment(dataset.transform_map)
    point_cloud, range_image, _ = dataset.load_range_image_points_from_file(os.path.join(REF, "assets/example_data/example.bin"))
    seg_idx, cidx, _ = oracle.segment(range_image[..., 0], dataset.transform_map, GROUND, cfg["cluster_num"])
    seg_idx = seg_idx.astype(np.int64)
    ground_model = GROUND.astype(np.float32)   # segment() returns the f32 tensor's numpy (utils/segment_utils.py:147)
    cluster_models = pc_seg.cluster_modeling(point_cloud, range_image, seg_idx, {"model_method": "point", "angle_threshold": 75})
    model_param = np.concatenate((ground_model.reshape(1, 4), cluster_models), 0)
    pred = pc_seg.intra_predict(seg_idx, model_param)
    residual = range_image - pred
    if nonuniform:
        QM = QuantizationModule(accuracy, uniform=False, level_kp_num=tuple(cfg["level_key_point_num"]),
                                level_dacc=tuple(cfg["level_delta_acc"]), ground_salience_level=cfg["ground_salience_level"],
                                feature_region=cfg["feature_region"], segments=cfg["segments"], sharp_num=cfg["sharp_num"],
                                less_sharp_num=cfg["less_sharp_num"], flat_num=cfg["flat_num"])
    else:
        QM = QuantizationModule(accuracy)
    rq, sal, kp = QM.quantize_residual(residual, seg_idx, point_cloud, range_image)
    original, compressed = compress_point_cloud(bc, model_param, seg_idx, sal, rq, point_cloud, range_image, full=False)
    with tempfile.NamedTemporaryFile(suffix=".rpcc") as f:
        save_compressed_bitstream(f.name, compressed, uniform=not nonuniform)
        blob = open(f.name, "rb").read()
        comp2 = read_compressed_bitstream(f.name, uniform=not nonuniform)
    H, W = dataset.transform_map.shape[:2]
    rq2, seg2, sal2, pp2 = decompress_point_cloud(comp2, bc, model_param.shape[0], H, W)
    res2 = (QuantizationModule(accuracy, uniform=True) if not nonuniform else QM).dequantize_residual(rq2, seg2, sal2)
    rec = pc_seg.intra_predict(seg2, pp2) + res2
    xyz = dataset.PCTransformer.range_image_to_point_cloud(rec)
    out = {
        "range_sha": sha(range_image.astype(np.float32)), "valid": int((range_image != 0).sum()),
        "seg_u8": seg_idx.astype(np.uint8), "center_idx": cidx,
        "model_param": model_param.astype(np.float32), "pred_sha": sha(pred.astype(np.float32)),
        "symbols": np.asarray(rq).astype(np.int16), "contour_bits": np.frombuffer(original["contour_map"].tobytes(), np.uint8),
        "idx_sequence": original["idx_sequence"], "rpcc": np.frombuffer(blob, np.uint8),
        "rec_sha": sha(rec.astype(np.float32)), "xyz_sha": sha(xyz.astype(np.float32)),
        "max_err": float(np.abs(rec - range_image).max()),
    }
    if nonuniform:
        out["key_points_u8"] = kp.astype(np.uint8)
        out["salience"] = np.asarray(sal).astype(np.uint8)
    return out


if __name__ == "__main__":
    # Reference bug: extract_features_with_segment never zero-fills key_point_map
    # (cpp_modules.cpp:38-43), so unwritten pixels hold whatever malloc returns and the salience
    # levels depend on heap history.  glibc's MALLOC_PERTURB_=255 zero-fills every malloc'd block,
    # which gives the only deterministic reading (fresh zero pages) -- re-exec under it.
    if os.environ.get("MALLOC_PERTURB_") != "255":
        os.environ["MALLOC_PERTURB_"] = "255"
        os.execv(sys.executable, [sys.executable] + sys.argv)
    g = {}
    for tag, nu in (("u", False), ("n", True)):
        for k, v in run(nu).items():
            g[tag + "_" + k] = v
    g["ground"] = GROUND
    # the only known-answer vector in the reference: the docstring of ContourExtractor.extract_contour
    # (utils/contour_utils.py:182-195)
    g["kat_idx_map"] = np.array([[1, 1, 1, 1, 2], [3, 2, 2, 1, 2], [3, 2, 1, 1, 2], [3, 3, 2, 2, 2]])
    g["kat_contour"] = np.array([[1, 0, 0, 0, 1], [1, 1, 0, 1, 1], [1, 1, 1, 0, 1], [1, 0, 1, 0, 0]])
    g["kat_seq"] = np.array([1, 2, 3, 2, 1, 2, 3, 2, 1, 2, 3, 2])
    np.savez_compressed(os.path.join(HERE, "example_golden.npz"), **g)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in g.items()})
    print("bytes", os.path.getsize(os.path.join(HERE, "example_golden.npz")))
