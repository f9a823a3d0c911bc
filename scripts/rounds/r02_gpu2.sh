#!/bin/bash
# round 2, visit 2: staged quantise kernel (cp.async.bulk + mbarrier) -- parity, then stage times with / without staging
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_stages.py tests/test_gpu_plane.py tests/test_gpu_datalist.py -m gpu -x -q 2>&1 | tail -4
echo "== staged"; python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "quantize|total"
echo "== direct"; RPCC_NO_STAGING=1 python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E "quantize|total"
echo "== staged nonuniform"; python scripts/stage_times.py 1184 10 nonuniform | tr ' ' '\n' | grep -E "quantize|keypoints|total"
echo "== staged plane"; python scripts/stage_times.py 1184 10 uniform plane | tr ' ' '\n' | grep -E "quantize|model|total"
