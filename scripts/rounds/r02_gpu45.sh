#!/bin/bash
# round 2, visit 45: FPS rounds hand the frames out longest first (live buckets counted by the first pass)
exec > gpurun_out/r02m_visit45.txt 2>&1
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
for o in 1 0 1 0; do
  echo "== RPCC_FPS_ORDER=$o: $(RPCC_FPS_ORDER=$o python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^fps|total' | tr '\n' ' ')"
done
for o in 1 0; do
  echo "== 2368 frames, RPCC_FPS_ORDER=$o: $(RPCC_FPS_ORDER=$o python scripts/stage_times.py 2368 6 | tr ' ' '\n' | grep -E '^fps|total' | tr '\n' ' ')"
done
