set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01d_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r01d_pytest.log
timeout 600 python bench.py > gpurun_out/r01d_bench.json 2> gpurun_out/r01d_bench.err; echo "bench exit $?"; cat gpurun_out/r01d_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r01d_bench_ref.json 2> gpurun_out/r01d_bench_ref.err; cat gpurun_out/r01d_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01d_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01d_launches.log 2>&1
tail -3 gpurun_out/r01d_launches.csv
bash scripts/ncu_full.sh r01d project_kernel segment_fps assign_labels_kernel quantize_pack_kernel ground_fit
ls -la gpurun_out
