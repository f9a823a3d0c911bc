#!/bin/bash
# round 2, visit 1: quantise kernel rewrite -- parity tests, then stage times of the variants
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_stages.py tests/test_gpu_plane.py -m gpu -x -q 2>&1 | tail -4
for v in "" q3 q5 q6 q4s2 q6s2 HEAD; do
  if [ -n "$v" ]; then export RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_$v.so; else unset RPCC_B200_LIB; fi
  echo "== ${v:-tree}"; python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "quantize|total"
done
