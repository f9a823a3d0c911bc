#!/bin/bash
# round 2, visit 9: FPS first pass (dead buckets skipped, shifted-coordinate boxes) and label statistics by match groups
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_plane.py -m gpu -x -q 2>&1 | tail -2
python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "fps|assign|total"
RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_HEAD.so python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "fps|assign|total"
