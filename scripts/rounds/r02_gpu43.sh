#!/bin/bash
# round 2, visit 43 (two GPUs): frames per pipeline stage of encode_host at N=2: 111 vs 148 (and 74, 222), same box
exec > gpurun_out/r02l_visit43.txt 2>&1
for hc in 111 148 74 222 111 148; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 2 --steps 10 --datalist-frames 0 --no-cpu-baseline --host-chunk $hc 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('host-chunk $hc: e2e', round(d['e2e']['value']), 'link ceiling', round(d['e2e']['link_ceiling_frames_per_s_all_gpus']), 'device', round(d['value']))"
done
