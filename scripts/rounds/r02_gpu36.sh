#!/bin/bash
# round 2, visit 36: wide FPS kernel with the maximum's point per bucket in shared memory: one barrier per round, no global load before the box tests
exec > gpurun_out/r02j_visit36.txt 2>&1
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
RPCC_FPS_WIDE=0 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "fps or segment" 2>&1 | tail -1
for i in 1 2; do echo "== $(python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps|total' | tr '\n' ' ')"; done
echo "== 32E: $(python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from rpcc_b200 import synthetic
from rpcc_b200.batch import BatchEncoder
for lidar in ("Velodyne32E", "VelodyneVLP16"):
    B = 1184
    per = [synthetic.frame(i, lidar) for i in range(16)]
    pts = np.concatenate([per[i % 16][0] for i in range(B)], 0)
    off = np.cumsum([0] + [per[i % 16][0].shape[0] for i in range(B)]).astype(np.int64)
    enc = BatchEncoder(lidar, accuracy=0.02, max_batch=B, max_points=pts.shape[0])
    d_pts, d_off = torch.from_numpy(pts).cuda(), torch.from_numpy(off).cuda()
    for r in range(3): enc.encode_device(0, d_pts, d_off, B, None)
    enc.sync(); enc.profile(True)
    for r in range(10): enc.encode_device(0, d_pts, d_off, B, None)
    ms, frames, calls = enc.stage_times()
    print(lidar, " ".join("%s=%.3f" % (k, v / calls) for k, v in ms.items() if k in ("fps", "assign", "project", "quantize")), end="; ")
PY
)"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:segment_fps_wide -c 1 --csv --log-file gpurun_out/r02j_fpswide2.csv python scripts/stage_times.py 1184 1 > /dev/null 2>&1
cut -d, -f13- gpurun_out/r02j_fpswide2.csv | tail -3 | rev | cut -d, -f1-3 | rev
