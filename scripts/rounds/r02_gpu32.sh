#!/bin/bash
# round 2, visit 32 (eight GPUs): bench at N=8 with each rank bound to the CPUs next to its GPU, and without (A/B of the e2e leg)
exec > gpurun_out/r02i_visit32.txt 2>&1
nvidia-smi topo -m | head -14
for g in 0 4; do cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $g | tr 'A-Z' 'a-z' | cut -c5-)/local_cpulist; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --datalist-frames 0 --no-cpu-baseline > gpurun_out/r02i_bench_n8_numa.json 2> gpurun_out/r02i_bench_n8_numa.err; tail -2 gpurun_out/r02i_bench_n8_numa.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 8 --datalist-frames 0 --no-cpu-baseline --no-numa-bind > gpurun_out/r02i_bench_n8_nobind.json 2> gpurun_out/r02i_bench_n8_nobind.err; tail -2 gpurun_out/r02i_bench_n8_nobind.err
