#!/bin/bash
# round 2, visit 31 (eight GPUs): bench as the driver launches it at N=8 (device / e2e / datalist / decode legs)
exec > gpurun_out/r02i_visit31.txt 2>&1
nvidia-smi -L | wc -l; nproc; free -g | head -2
date +%s
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 > gpurun_out/r02i_bench_n8.json 2> gpurun_out/r02i_bench_n8.err; tail -3 gpurun_out/r02i_bench_n8.err
date +%s
