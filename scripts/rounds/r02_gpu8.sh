#!/bin/bash
# round 2, visit 8: hierarchical FPS kernel with packed bucket records and ILP-way updates
python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "fps or segment" 2>&1 | tail -2
for v in "" ilp1 ilp3 ilp4; do
  if [ -n "$v" ]; then export RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_$v.so; else unset RPCC_B200_LIB; fi
  for t in 1024 768 512; do echo "== ${v:-ilp2} threads $t: $(RPCC_FPS_THREADS=$t python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps')"; done
done
