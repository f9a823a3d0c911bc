# assign<MQ> + ground 512x2: tests, stage times, and the device-side bench at several frames-per-launch.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python scripts/stage_times.py 296 10 2>&1 | tail -1
python scripts/stage_times.py 592 10 2>&1 | tail -1
for mb in 148 296 592 1184; do
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --max-batch $mb --frames $((4*mb)) --steps 10 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('max_batch', j['config']['frames_per_launch'], 'value', round(j['value']), 'ms/step', round(j['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in j['roofline']['kernels'].items()})"
done
