#!/bin/bash
# round 2, visit 20: FPS rounds kernel (1 CTA / SM, batched bucket loads, winner coordinates in shared memory) vs the 2-CTA one
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
for nb in 4 2 6 0; do
  echo "== RPCC_FPS_NBATCH=$nb"; RPCC_FPS_NBATCH=$nb python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "^fps|total"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02g_fps_launches.csv python scripts/stage_times.py 1184 2 > /dev/null 2>&1
grep -E "fps" gpurun_out/r02g_fps_launches.csv | tail -2 | cut -d, -f5,15
