#!/usr/bin/env python
"""A/B check of the ground-fit kernel: the fitted planes of N synthetic frames (per lidar) must be
bit-identical between two builds of the library.
   python scripts/ab_ground.py dump out.npz          (run once per library, RPCC_B200_LIB selects it)
   python scripts/ab_ground.py cmp a.npz b.npz"""
import sys

import numpy as np


def dump(path):
    import torch
    from rpcc_b200 import synthetic
    from rpcc_b200.batch import BatchEncoder
    out = {}
    for lidar in ("Velodyne64E", "Velodyne32E", "VelodyneVLP16"):
        seeds = list(range(300, 324))
        pts, off, _ = synthetic.batch(seeds, lidar)
        d_pts, d_off = torch.from_numpy(pts).cuda(), torch.from_numpy(off).cuda()
        method = __import__("os").environ.get("AB_METHOD", "point")      # AB_METHOD=plane: per-cluster plane models too
        with BatchEncoder(lidar, accuracy=0.02, max_batch=len(seeds), model_method=method) as enc:
            enc.encode_device(0, d_pts, d_off, len(seeds), None)
            enc.sync()
            out[lidar] = enc.device_buffer(0, "ground", (len(seeds), 4), torch.float32).cpu().numpy().copy()
            out[lidar + "_model"] = enc.device_buffer(0, "model", (len(seeds), 102, 4), torch.float32).cpu().numpy().copy()
            sb = enc.device_buffer(0, "sym_base", (len(seeds) + 1,), torch.int64).cpu().numpy()
            out[lidar + "_sym"] = enc.device_buffer(0, "symbols", (int(sb[-1]),), torch.int16).cpu().numpy().copy()
    np.savez(path, **out)


def cmp(a, b):
    A, B = np.load(a), np.load(b)
    ok = True
    for k in A.files:
        same = A[k].shape == B[k].shape and np.array_equal(A[k].view(np.uint8), B[k].view(np.uint8))
        print(k, "identical" if same else "DIFFERENT")
        ok &= same
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    if sys.argv[1] == "dump":
        dump(sys.argv[2])
    else:
        cmp(sys.argv[2], sys.argv[3])
