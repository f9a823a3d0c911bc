"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): 3 frames per framework through the
encoder (device ground fit, the staged quantise kernel, the eval stage), the decoder (packed streams, row compaction)
and chamfer.  python scripts/sanitize_small.py [lidar ...]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rpcc_b200 import synthetic  # noqa: E402
from rpcc_b200.batch import BatchDecoder, BatchEncoder  # noqa: E402
from rpcc_b200.evaluate_metrics import calc_chamfer_distance  # noqa: E402

cases = (("Velodyne64E", False, "point"), ("Velodyne64E", True, "point"), ("Velodyne64E", False, "plane"),
         ("VelodyneVLP16", True, "point"), ("VelodyneVLP16", False, "point"))
only = set(sys.argv[1:])
for lidar, nonuniform, method in cases:
    if only and lidar not in only:
        continue
    pts, off, grounds = synthetic.batch([70, 71, 72], lidar)
    with BatchEncoder(lidar, accuracy=0.02, nonuniform=nonuniform, max_batch=3, model_method=method, eval=True) as enc:
        blobs = enc.compress(pts, off, None)          # ground fitted on the device
    d = BatchDecoder(lidar, accuracy=0.02, nonuniform=nonuniform).decode(blobs, want_points=True)
    xyz = d["xyz"][0].reshape(-1, 3).cpu().numpy()
    r = calc_chamfer_distance(pts[off[0]:off[1], :3][::8], xyz[::8], out=False)
    print(lidar, nonuniform, method, [len(b) for b in blobs], "chamfer mean %.4f" % r["mean"])
# the FPS round kernel's tie path (several buckets hold the frame's maximum): a sphere, nothing masked
from rpcc_b200 import LidarConfig, device  # noqa: E402
cfg = LidarConfig("VelodyneVLP16")
lut = torch.from_numpy(cfg.transform_map()).cuda()
sphere = torch.full((2, cfg.H, cfg.W), 10.0, device="cuda")
cidx, _ = device.segment_fps_batch(sphere, lut, torch.tensor([[0.0, 0.0, 1.0, 100.0]] * 2, device="cuda"), 100, 0.1)
print("tie path seeds", cidx[0, :6].tolist())
torch.cuda.synchronize()
print("done")
