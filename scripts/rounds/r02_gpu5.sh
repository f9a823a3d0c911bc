#!/bin/bash
# round 2, visit 5 (2 GPUs): the two-rank torchrun test of the batch tool, then bench.py under torchrun at N=2
nvidia-smi -L
python -m pytest tests/test_gpu_datalist.py -m gpu -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
tail -3 gpurun_out/r02d_bench_n2.err; cat gpurun_out/r02d_bench_n2.json | cut -c1-300
