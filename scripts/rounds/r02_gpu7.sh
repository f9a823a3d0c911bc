#!/bin/bash
# round 2, visit 7: hierarchical FPS kernel -- seed parity vs the reference kernel, stage times per thread count, old kernel beside
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
for t in 1024 768 512; do echo "== hier $t"; RPCC_FPS_THREADS=$t python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "fps|total"; done
echo "== pruned (round 1)"; RPCC_FPS_IMPL=pruned python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "fps|total"
