# FPS frame queue: bit-for-bit against HEAD, stage times, device bench at 296 / 1184 frames per launch.
mkdir -p /tmp/ab
AB=r-pcc_b200/build/ab
RPCC_B200_LIB=$PWD/$AB/librpcc_HEAD.so python scripts/ab_ground.py dump /tmp/ab/HEAD.npz 2>&1 | tail -1
python scripts/ab_ground.py dump /tmp/ab/cur.npz 2>&1 | tail -1
python scripts/ab_ground.py cmp /tmp/ab/HEAD.npz /tmp/ab/cur.npz
python scripts/stage_times.py 296 10 2>&1 | tail -1
for mb in 296 1184; do
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --max-batch $mb --frames $((4*mb)) --steps 10 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('max_batch', j['config']['frames_per_launch'], 'value', round(j['value']), 'ms/step', round(j['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in j['roofline']['kernels'].items()})"
done
