#!/bin/bash
# round 2, visit 35: FPS rounds with 64-pixel buckets: 4 CTAs of 512 threads / 2 CTAs of 1024 threads per SM, against the pair kernel
exec > gpurun_out/r02j_visit35.txt 2>&1
for w in 512 1024; do RPCC_FPS_WIDE=$w python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "fps or segment" 2>&1 | tail -1; done
for w in 512 1024 0 512 0; do
  echo "== RPCC_FPS_WIDE=$w: $(RPCC_FPS_WIDE=$w python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps|total' | tr '\n' ' ')"
done
RPCC_FPS_WIDE=512 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:segment_fps_wide -c 1 --csv --log-file gpurun_out/r02j_fpswide.csv python scripts/stage_times.py 1184 1 > /dev/null 2>&1
cut -d, -f13- gpurun_out/r02j_fpswide.csv | tail -3
