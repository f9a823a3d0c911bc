#!/bin/bash
# round 2, visit 30: assign -- square root only when some lane can still improve
exec > gpurun_out/r02h_visit30.txt 2>&1
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do echo "== $(python scripts/stage_times.py 1184 10 | tr '\n' ' ')"; done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:assign_labels -c 2 --csv --log-file gpurun_out/r02h_assign_launches.csv python scripts/stage_times.py 1184 1 > /dev/null 2>&1
cut -d, -f5,13- gpurun_out/r02h_assign_launches.csv | tail -2
