#!/bin/bash
# round 2, visit 42 (two GPUs): bench at N=2 on the final tree
exec > gpurun_out/r02l_visit42.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus 2 > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err; tail -2 gpurun_out/r02l_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02l_bench_n2.json') if l.startswith('{')][0])
print('N', d['n_gpus'], 'device', round(d['value']), 'e2e', round(d['e2e']['value']), 'ceiling', round(d['e2e']['link_ceiling_frames_per_s_all_gpus']), 'datalist', round(d['e2e']['datalist']['value']), d['e2e']['datalist']['consistency'], 'decode', round(d['e2e']['decode']['value']))
PY
