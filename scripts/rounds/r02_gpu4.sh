#!/bin/bash
# round 2, visit 4: quantise kernel -- contour words stored once per tile; full parity suite
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== staged"; python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "quantize|total"
echo "== direct"; RPCC_NO_STAGING=1 python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "quantize|total"
echo "== direct occ5"; RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_occ5.so RPCC_NO_STAGING=1 python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "quantize|total"
for l in Velodyne32E VelodyneVLP16; do echo "== $l"; python - <<PY
import sys; sys.path.insert(0,'.')
import numpy as np, torch
from rpcc_b200 import synthetic
from rpcc_b200.batch import BatchEncoder
B=1184
per=[synthetic.frame(i,"$l") for i in range(16)]
pts=np.concatenate([per[i%16][0] for i in range(B)],0); off=np.cumsum([0]+[per[i%16][0].shape[0] for i in range(B)]).astype(np.int64)
enc=BatchEncoder("$l",accuracy=0.02,max_batch=B,max_points=pts.shape[0])
d_pts,d_off=torch.from_numpy(pts).cuda(),torch.from_numpy(off).cuda()
for r in range(3): enc.encode_device(0,d_pts,d_off,B,None)
enc.sync(); enc.profile(True)
for r in range(10): enc.encode_device(0,d_pts,d_off,B,None)
ms,fr,calls=enc.stage_times()
print(" ".join("%s=%.4f"%(k,v/calls) for k,v in ms.items()))
PY
done
