#!/bin/bash
# round 2, visit 26: pair FPS kernel with the update split by seed-0 case; quantise: plane flag in the sign of 1/step, contour word reversed once per tile
exec > gpurun_out/r02h_visit26.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
  echo "== $(python scripts/stage_times.py 1184 10 | tr '\n' ' ')"
done
