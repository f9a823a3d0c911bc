#!/bin/bash
# round 2, visit 46: rank kernel of the FPS frame queue through shared memory
exec > gpurun_out/r02m_visit46.txt 2>&1
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -1
for i in 1 2; do echo "== $(python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps|total' | tr '\n' ' ')"; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fps_order -c 2 --csv --log-file gpurun_out/r02m_order.csv python scripts/stage_times.py 1184 1 > /dev/null 2>&1
tail -1 gpurun_out/r02m_order.csv | rev | cut -d, -f1-3 | rev
