#!/bin/bash
# round 2, visit 52: the committed tree at the end of the round: GPU suite, smoke, bench
exec > gpurun_out/r02o_visit52.txt 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; tail -1 gpurun_out/r02o_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02o_bench.json') if l.startswith('{')][0])
print('device', round(d['value']), 'e2e', round(d['e2e']['value']), 'ceiling', round(d['e2e']['link_ceiling_frames_per_s_all_gpus']), 'datalist', round(d['e2e']['datalist']['value']), d['e2e']['datalist']['consistency'], 'decode', round(d['e2e']['decode']['value']), 'launches', d['gpu_launches'])
PY
