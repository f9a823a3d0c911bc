# one GPU-box visit: tests, bench (both arms), stage times, launch list, ncu --set full of the hot kernels (1184 frames per launch).
TAG=${1:-r01f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json
python scripts/stage_times.py 296 20 | tee gpurun_out/${TAG}_stage_times.txt
python scripts/stage_times.py 1184 10 | tee -a gpurun_out/${TAG}_stage_times.txt
python scripts/stage_times.py 1184 10 nonuniform | tee -a gpurun_out/${TAG}_stage_times.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
FRAMES=1184 REPS=2 bash scripts/ncu_full.sh ${TAG} project_kernel segment_fps assign_labels_kernel quantize_pack_kernel ground_fit_kernel
ls -la gpurun_out
