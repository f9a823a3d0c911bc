#!/bin/bash
# round 2, visit 53: the last host-coder commits on the box: datalist / pipeline / drop-in tests
exec > gpurun_out/r02o_visit53.txt 2>&1
python -m pytest tests/test_gpu_datalist.py tests/test_gpu_pipeline.py tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -2
