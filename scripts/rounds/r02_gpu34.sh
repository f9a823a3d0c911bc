#!/bin/bash
# round 2, visit 34: pair FPS kernel with 768 threads (40 registers, 24 warps, Q = 6) against 1024 threads
exec > gpurun_out/r02j_visit34.txt 2>&1
RPCC_FPS_THREADS=768 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "fps or segment" 2>&1 | tail -2
for t in 768 1024 768 1024; do
  echo "== RPCC_FPS_THREADS=$t: $(RPCC_FPS_THREADS=$t python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps|total' | tr '\n' ' ')"
done
RPCC_FPS_THREADS=768 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:segment_fps_pair -c 1 --csv --log-file gpurun_out/r02j_fps768.csv python scripts/stage_times.py 1184 1 > /dev/null 2>&1
cut -d, -f13- gpurun_out/r02j_fps768.csv | tail -3
