#!/bin/bash
# round 2, visit 38: parity off the default configuration (accuracy 0.005..0.05, 20..250 clusters)
exec > gpurun_out/r02l_visit38.txt 2>&1
python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "other_accuracies" 2>&1 | tail -15
