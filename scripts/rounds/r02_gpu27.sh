#!/bin/bash
# round 2, visit 27: new FPS tests (odd image, ties), compute-sanitizer over the pair FPS kernel incl. its tie path, timed bench (both arms)
exec > gpurun_out/r02h_visit27.txt 2>&1
python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "fps or segment" 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize_small.py VelodyneVLP16 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python scripts/sanitize_small.py VelodyneVLP16 2>&1 | tail -6
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 1 python scripts/sanitize_small.py VelodyneVLP16 2>&1 | tail -4
date +%s.%N
python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -2 gpurun_out/r02h_bench.err
date +%s.%N
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02h_bench_ref.json 2> gpurun_out/r02h_bench_ref.err; tail -2 gpurun_out/r02h_bench_ref.err
date +%s.%N
