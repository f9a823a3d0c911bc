#!/bin/bash
# round 2, visit 41: two-pixel label kernel at 6 and 8 CTAs per SM
exec > gpurun_out/r02l_visit41.txt 2>&1
for v in as2o6 as2o8 as2o6 as2o8; do
  echo "== $v: $(RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_$v.so python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^assign|total' | tr '\n' ' ')"
done
RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_as2o8.so python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -1
