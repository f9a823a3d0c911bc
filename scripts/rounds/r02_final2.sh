#!/bin/bash
# round 2, final visit of the second session (one B200): GPU suite, smoke, both bench arms, launch list, ncu --set full of the hot kernels
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02i_pytest_gpu.log; cat gpurun_out/r02i_pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -2 gpurun_out/r02i_bench.err
python bench.py --impl reference > gpurun_out/r02i_bench_ref.json 2> gpurun_out/r02i_bench_ref.err; tail -2 gpurun_out/r02i_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 2 --warmup 1 --passes 1 --no-e2e --no-cpu-baseline > gpurun_out/r02i_launches.log 2>&1
FRAMES=1184 REPS=4 bash scripts/ncu_full.sh r02i project_kernel fps_first_pass4_kernel segment_fps_pair_kernel assign_labels_kernel quantize_pack_staged_kernel ground_fit_kernel
python scripts/stage_times.py 1184 10 > gpurun_out/r02i_stage_times.txt; python scripts/stage_times.py 1184 10 nonuniform >> gpurun_out/r02i_stage_times.txt; python scripts/stage_times.py 1184 10 uniform plane >> gpurun_out/r02i_stage_times.txt; cat gpurun_out/r02i_stage_times.txt
python scripts/bench_rows.py 296 > gpurun_out/r02i_rows.json 2> gpurun_out/r02i_rows.err; tail -3 gpurun_out/r02i_rows.err
