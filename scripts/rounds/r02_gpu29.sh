#!/bin/bash
# round 2, visit 29: quantise -- branch-free plane prediction (tree), 6 warps per CTA at 5 / 6 CTAs per SM, no __syncwarp around the counters
exec > gpurun_out/r02h_visit29.txt 2>&1
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_plane.py -m gpu -x -q 2>&1 | tail -2
echo "== tree: $(python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^quantize' | tr '\n' ' ')"
for v in w6 w6o5 nosync w6nosync; do
  echo "== $v: $(RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_$v.so python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^quantize' | tr '\n' ' ')"
done
echo "== tree: $(python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^quantize' | tr '\n' ' ')"
