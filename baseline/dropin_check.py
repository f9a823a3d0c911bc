#!/usr/bin/env python
"""The reference's UNMODIFIED L3 Python (baseline/_ref/R-PCC: dataset/, utils/compress_utils.py, utils/segment_utils.py,
utils/evaluate_metrics.py) run on librpcc_b200.so through the `ops` shims of baseline/ops_b200 -- BASELINE configs[1] as
one piece: non-uniform encode -> .rpcc -> decode -> chamfer.  Prints one JSON line of digests and figures;
tests/test_gpu_dropin.py compares it with tests/golden/example_golden.npz (made by the same Python on the reference's
own C++).   python baseline/dropin_check.py <input.bin> <uniform|nonuniform> [reference|b200]"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_reference_tool  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    path, mode = sys.argv[1], sys.argv[2]
    flavour = sys.argv[3] if len(sys.argv) > 3 else "b200"
    nonuniform = mode == "nonuniform"
    ref = run_reference_tool.setup_paths(flavour)
    from dataset import build_dataset
    from utils.compress_utils import (BasicCompressor, QuantizationModule, compress_point_cloud, decompress_point_cloud,
                                      read_compressed_bitstream, save_compressed_bitstream)
    from utils.evaluate_metrics import calc_chamfer_distance
    from utils.segment_utils import PointCloudSegment
    from utils.utils import load_compressor_cfg
    yaml = os.path.join(ref, "cfgs", "compressor.yaml")
    cfg = load_compressor_cfg(yaml)
    accuracy = cfg["accuracy"] * 2
    bc = BasicCompressor(compressor_yaml=yaml)
    dataset = build_dataset(lidar_type="Velodyne64E")
    pc_seg = PointCloudSegment(dataset.transform_map)
    point_cloud, range_image, _ = dataset.load_range_image_points_from_file(path)
    segment_cfg = {"segment_method": "FPS", "ground_vertical_threshold": cfg["ground_threshold"], "cluster_num": cfg["cluster_num"],
                   "DBSCAN_eps": cfg["DBSCAN_eps"]}
    seg_idx, ground_model = pc_seg.segment(point_cloud, range_image, segment_cfg, cpu=False)     # torch ops + FPS plugin
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
    QM2 = QM if nonuniform else QuantizationModule(accuracy, uniform=True)
    res2 = QM2.dequantize_residual(rq2, seg2, sal2)
    rec = pc_seg.intra_predict(seg2, pp2) + res2
    xyz = dataset.PCTransformer.range_image_to_point_cloud(rec)
    ch = calc_chamfer_distance(point_cloud, xyz, out=False)
    out = {"range_sha": sha(range_image.astype(np.float32)), "seg_sha": sha(seg_idx.astype(np.uint8)),
           "model_sha": sha(model_param.astype(np.float32)), "pred_sha": sha(pred.astype(np.float32)),
           "symbols_sha": sha(np.asarray(rq).astype(np.int16)), "rpcc_sha": sha(np.frombuffer(blob, np.uint8)),
           "rpcc_bytes": len(blob), "rec_sha": sha(rec.astype(np.float32)), "xyz_sha": sha(xyz.astype(np.float32)),
           "max_err": float(np.abs(rec - range_image).max()), "chamfer_mean": ch["mean"], "f_score": ch["f_score"],
           "cd1": ch["cd1"], "cd2": ch["cd2"], "dist1_sha": sha(ch["chamfer_dist_info"]["dist1"]),
           "flavour": flavour, "segment_utils": sys.modules["utils.segment_utils"].__file__}
    if nonuniform:
        out["salience_sha"] = sha(np.asarray(sal).astype(np.uint8))
        out["key_points_sha"] = sha(kp.astype(np.uint8))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
