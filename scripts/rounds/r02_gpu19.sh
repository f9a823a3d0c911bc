#!/bin/bash
# round 2, visit 19: FPS -- round 1 (distances, boxes, maxima) in its own streaming kernel vs inside the round kernel
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
echo "== split"; python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "^fps|total"
echo "== fused"; RPCC_FPS_SPLIT=0 python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "^fps|total"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02g_fps_launches.csv python scripts/stage_times.py 1184 2 > /dev/null 2>&1
grep -E "fps" gpurun_out/r02g_fps_launches.csv | tail -4
