#!/bin/bash
# round 2, visit 44: FPS seeds of 64 + 48 distinct frames against the reference kernel
exec > gpurun_out/r02l_visit44.txt 2>&1
python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "many_frames_against" 2>&1 | tail -12
RPCC_FPS_WIDE=0 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "many_frames_against" 2>&1 | tail -2
