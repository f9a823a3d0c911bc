#!/bin/bash
# round 2, visit 13: degenerate frames; compute-sanitizer over the new kernels (memcheck all, racecheck + synccheck on VLP16)
python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "degenerate" 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize_small.py 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python scripts/sanitize_small.py VelodyneVLP16 2>&1 | tail -4
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 1 python scripts/sanitize_small.py VelodyneVLP16 2>&1 | tail -4
