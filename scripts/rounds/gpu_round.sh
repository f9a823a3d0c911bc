# one GPU-box visit: tests, bench (both arms), PCIe probe, e2e chunk sweep.  TAG names the outputs.
TAG=${1:-r01e}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv
nproc; free -g | head -2
timeout 300 python scripts/pcie_probe.py > gpurun_out/${TAG}_pcie.json 2> gpurun_out/${TAG}_pcie.err; cat gpurun_out/${TAG}_pcie.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json
for hc in 37 74 296; do
  timeout 300 python bench.py --steps 5 --no-cpu-baseline --host-chunk $hc > gpurun_out/${TAG}_bench_hc$hc.json 2> gpurun_out/${TAG}_bench_hc$hc.err
  python -c "import json,sys; d=json.load(open('gpurun_out/${TAG}_bench_hc$hc.json')); print($hc, d['value'], d['e2e'])"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json
ls -la gpurun_out
