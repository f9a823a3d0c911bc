#!/bin/bash
# round 2, visit 17: label kernel -- sphere centre from shifted coordinates; FPS vector boxes; parity suite
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_plane.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -2
echo "== tree"; python scripts/stage_times.py 1184 10
echo "== HEAD"; RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_HEAD.so python scripts/stage_times.py 1184 10
