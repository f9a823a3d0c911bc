#!/bin/bash
# round 2, visit 49 (four GPUs): bench at N=4 on the final tree
exec > gpurun_out/r02n_visit49.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 > gpurun_out/r02n_bench_n4.json 2> gpurun_out/r02n_bench_n4.err; tail -2 gpurun_out/r02n_bench_n4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02n_bench_n4.json') if l.startswith('{')][0])
print('N', d['n_gpus'], 'device', round(d['value']), 'e2e', round(d['e2e']['value']), 'ceiling', round(d['e2e']['link_ceiling_frames_per_s_all_gpus']), 'datalist', round(d['e2e']['datalist']['value']), d['e2e']['datalist']['consistency'], 'decode', round(d['e2e']['decode']['value']), 'cores', d['e2e']['datalist']['host_cores'])
PY
