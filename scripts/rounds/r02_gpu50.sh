#!/bin/bash
# round 2, visit 50: frames per launch of the device-resident leg: 1184 (default) vs 2368 vs 592
exec > gpurun_out/r02n_visit50.txt 2>&1
for mb in 1184 2368 592 1184 2368; do
  python bench.py --steps 10 --no-e2e --no-cpu-baseline --max-batch $mb 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('max-batch $mb: device', round(d['value']), 'ms/step', round(d['ms_per_step'],2))"
done
