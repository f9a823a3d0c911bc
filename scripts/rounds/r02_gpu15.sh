#!/bin/bash
# round 2, visit 15: FPS -- the round's winner reduced and fetched by warp 0 alone (two syncs) vs by every warp (one sync)
python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "fps or segment" 2>&1 | tail -2
echo "== one winner"; python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "^fps|total"
echo "== all warps"; RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_allwin.so python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "^fps|total"
