TAG=${1:-r01f}
set -x
mkdir -p gpurun_out
RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_HEAD.so python scripts/ab_ground.py dump gpurun_out/${TAG}_ground_prev.npz
python scripts/ab_ground.py dump gpurun_out/${TAG}_ground_new.npz
python scripts/ab_ground.py cmp gpurun_out/${TAG}_ground_prev.npz gpurun_out/${TAG}_ground_new.npz
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['h2d_achieved_gbs'])
print({k:round(v['ms_per_launch'],4) for k,v in d['roofline']['kernels'].items()})
PY
for hc in 37 74 148 296; do
  timeout 300 python bench.py --steps 5 --no-cpu-baseline --host-chunk $hc > gpurun_out/${TAG}_bench_hc$hc.json 2> gpurun_out/${TAG}_bench_hc$hc.err
  python -c "import json,sys; d=json.load(open('gpurun_out/${TAG}_bench_hc$hc.json')); print($hc, d['value'], d['e2e']['value'], d['e2e']['h2d_achieved_gbs'])"
done
