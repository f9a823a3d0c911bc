#!/bin/bash
# round 2, visit 6: quantise kernel, step loop unrolled 1 / 2 / 4
for v in "" u2 u4; do
  if [ -n "$v" ]; then export RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_$v.so; else unset RPCC_B200_LIB; fi
  echo "== ${v:-tree}"; python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "quantize|total"
done
