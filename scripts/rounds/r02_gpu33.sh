#!/bin/bash
# round 2, visit 33: quantise -- plane prediction by reciprocal product with an exact fallback, one slice counter, 64-bit table reads
exec > gpurun_out/r02j_visit33.txt 2>&1
python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_plane.py tests/test_gpu_datalist.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do echo "== $(python scripts/stage_times.py 1184 10 | tr ' ' '\n' | grep -E '^quantize|total' | tr '\n' ' ')"; done
echo "== plane: $(python scripts/stage_times.py 1184 10 uniform plane | tr ' ' '\n' | grep -E '^quantize|total' | tr '\n' ' ')"
echo "== nonuniform: $(python scripts/stage_times.py 1184 10 nonuniform | tr ' ' '\n' | grep -E '^quantize|total' | tr '\n' ' ')"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:quantize_pack -c 2 --csv --log-file gpurun_out/r02j_quant_launches.csv python scripts/stage_times.py 1184 1 > /dev/null 2>&1
cut -d, -f5,13- gpurun_out/r02j_quant_launches.csv | tail -2
