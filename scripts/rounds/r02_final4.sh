#!/bin/bash
# round 2, last visit (one B200): GPU suite, smoke, both bench arms, launch list, ncu --set full of the hot kernels, e2e chunk sweep
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02m_pytest_gpu.log; cat gpurun_out/r02m_pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
python bench.py > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; tail -2 gpurun_out/r02m_bench.err
python bench.py --impl reference > gpurun_out/r02m_bench_ref.json 2> gpurun_out/r02m_bench_ref.err; tail -2 gpurun_out/r02m_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches.csv python bench.py --steps 2 --warmup 1 --passes 1 --no-e2e --no-cpu-baseline > gpurun_out/r02m_launches.log 2>&1
FRAMES=1184 REPS=4 bash scripts/ncu_full.sh r02m project_kernel fps_first_pass4_kernel segment_fps_wide_kernel assign_labels_kernel quantize_pack_staged_kernel ground_fit_kernel
python scripts/stage_times.py 1184 10 > gpurun_out/r02m_stage_times.txt; python scripts/stage_times.py 1184 10 nonuniform >> gpurun_out/r02m_stage_times.txt; python scripts/stage_times.py 1184 10 uniform plane >> gpurun_out/r02m_stage_times.txt; cat gpurun_out/r02m_stage_times.txt
for hc in; do
  python bench.py --steps 10 --datalist-frames 0 --no-cpu-baseline --host-chunk $hc 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('host-chunk $hc: e2e', round(d['e2e']['value']), 'link ceiling', round(d['e2e']['link_ceiling_frames_per_s_all_gpus']), 'device', round(d['value']))" | tee -a gpurun_out/r02m_host_chunk.txt
done
