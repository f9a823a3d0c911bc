#!/bin/bash
# round 2, visit 21: FPS -- rotation of the bucket -> warp map (balance of a round's bucket updates over the warps)
python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "fps or segment" 2>&1 | tail -1
for rot in 0 3 5 8 11 13 16; do
  echo "== 2-CTA kernel, RPCC_FPS_ROT=$rot: $(RPCC_FPS_NBATCH=0 RPCC_FPS_ROT=$rot python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps')"
done
for rot in 0 8 13; do
  echo "== rounds kernel, RPCC_FPS_ROT=$rot: $(RPCC_FPS_NBATCH=4 RPCC_FPS_ROT=$rot python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps')"
done
