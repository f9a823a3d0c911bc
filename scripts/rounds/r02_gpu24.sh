#!/bin/bash
# round 2, visit 24: FPS round kernel with two buckets per step (16 lanes x 2 pixels) and the winner found by a bucket scan
exec > gpurun_out/r02h_fps_pair.txt 2>&1
echo "== tests, pair kernel"; python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
echo "== tests, old kernel on the new tie test"; RPCC_FPS_PAIR=0 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "fps or segment" 2>&1 | tail -3
for pair in 1 0 1 0; do
  echo "== RPCC_FPS_PAIR=$pair: $(RPCC_FPS_PAIR=$pair python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps|total' | tr '\n' ' ')"
done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:fps -c 12 --csv --log-file gpurun_out/r02h_fps_launches.csv python scripts/stage_times.py 1184 2 > /dev/null 2>&1
cut -d, -f5,13- gpurun_out/r02h_fps_launches.csv | tail -8
