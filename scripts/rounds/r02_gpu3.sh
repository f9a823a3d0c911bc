#!/bin/bash
# round 2, visit 3: staged quantise kernel with the LUT requests hoisted above the rank / contour work; 4 vs 2 slices per step
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_stages.py -m gpu -x -q 2>&1 | tail -2
for v in "" s2; do
  if [ -n "$v" ]; then export RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_$v.so; else unset RPCC_B200_LIB; fi
  echo "== ${v:-tree}"; python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "quantize|total"
done
