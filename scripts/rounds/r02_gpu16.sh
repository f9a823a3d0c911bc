#!/bin/bash
# round 2, visit 16: FPS -- bucket boxes as float4 + float2 (two shared loads per test)
python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "fps or segment" 2>&1 | tail -2
echo "== vector boxes"; python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "^fps|total"
echo "== HEAD"; RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_HEAD.so python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "^fps|total"
