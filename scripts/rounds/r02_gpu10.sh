#!/bin/bash
# round 2, visit 10: FPS update fast path (seed 0 at the origin) + the general path under test
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "fps|assign|total"
