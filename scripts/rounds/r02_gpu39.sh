#!/bin/bash
# round 2, visit 39: ground fit with a 128-bit candidate scan; label kernel statistics by run scan + shared-memory atomics
exec > gpurun_out/r02l_visit39.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do echo "== $(python scripts/stage_times.py 1184 10 | tr '\n' ' ')"; done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"assign_labels|ground_fit" -c 4 --csv --log-file gpurun_out/r02l_ag.csv python scripts/stage_times.py 1184 1 > /dev/null 2>&1
grep -E "assign_labels|ground_fit" gpurun_out/r02l_ag.csv | tail -4 | rev | cut -d, -f1-3 | rev
