"""Runs the encoder chain a few times on ONE stream slot (so that a profiler sees each kernel in
isolation): python scripts/profile_chain.py [frames] [reps].  Used under ncu, see profiles/README.md."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rpcc_b200 import synthetic  # noqa: E402
from rpcc_b200.batch import BatchEncoder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nonuniform = len(sys.argv) > 3 and sys.argv[3] == "nonuniform"
per = [synthetic.frame(i) for i in range(min(B, 32))]
pts = np.concatenate([per[i % len(per)][0] for i in range(B)], 0)
off = np.cumsum([0] + [per[i % len(per)][0].shape[0] for i in range(B)]).astype(np.int64)
g = np.stack([per[i % len(per)][1] for i in range(B)]).astype(np.float32)
enc = BatchEncoder("Velodyne64E", accuracy=0.02, nonuniform=nonuniform, max_batch=B, max_points=pts.shape[0])
d_pts, d_off, d_g = torch.from_numpy(pts).cuda(), torch.from_numpy(off).cuda(), torch.from_numpy(g).cuda()
for r in range(reps):
    enc.encode_device(0, d_pts, d_off, B, None if r % 2 else d_g)
    enc.sync()
print("done", B, reps)
