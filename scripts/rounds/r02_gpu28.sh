#!/bin/bash
# round 2, visit 28 (two GPUs): the torchrun / NCCL datalist test, bench at N=2 as the driver launches it
exec > gpurun_out/r02h_visit28.txt 2>&1
nvidia-smi -L
python -m pytest tests/test_gpu_datalist.py -m gpu -x -q 2>&1 | tail -3
date +%s
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; tail -2 gpurun_out/r02h_bench_n2.err
date +%s
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_bench_ref_n2.json 2> gpurun_out/r02h_bench_ref_n2.err; tail -2 gpurun_out/r02h_bench_ref_n2.err
date +%s
