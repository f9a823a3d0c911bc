#!/bin/bash
# round 2, visit 40: label kernel with two pixels per lane (64-pixel slices)
exec > gpurun_out/r02l_visit40.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== one px/lane: $(RPCC_AS_PX=1 python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^assign|total' | tr '\n' ' ')"
echo "== two px/lane, 4 CTAs/SM: $(python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^assign|total' | tr '\n' ' ')"
for v in as2o3 as2o5 as2o6; do
  echo "== $v: $(RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_$v.so python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^assign|total' | tr '\n' ' ')"
done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:assign_labels -c 1 --csv --log-file gpurun_out/r02l_as2.csv python scripts/stage_times.py 1184 1 > /dev/null 2>&1
tail -3 gpurun_out/r02l_as2.csv | rev | cut -d, -f1-3 | rev
