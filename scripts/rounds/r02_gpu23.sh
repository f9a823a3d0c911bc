#!/bin/bash
# round 2, visit 23: repeat of visits 20/21 with the output kept (FPS rounds kernel vs 2-CTA kernel, bucket -> warp rotation)
exec > gpurun_out/r02h_fps_ab.txt 2>&1
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
for nb in 4 2 6 0; do
  echo "== RPCC_FPS_NBATCH=$nb ROT=8: $(RPCC_FPS_NBATCH=$nb python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps|total' | tr '\n' ' ')"
done
for rot in 0 3 5 8 13; do
  echo "== 2-CTA kernel, RPCC_FPS_ROT=$rot: $(RPCC_FPS_NBATCH=0 RPCC_FPS_ROT=$rot python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps' | tr '\n' ' ')"
done
for rot in 0 5 13; do
  echo "== rounds kernel NBATCH=4, RPCC_FPS_ROT=$rot: $(RPCC_FPS_NBATCH=4 RPCC_FPS_ROT=$rot python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps' | tr '\n' ' ')"
done
