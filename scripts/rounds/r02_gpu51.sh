#!/bin/bash
# round 2, visit 51: own bzip2 decoder in rpcc_unpack_rpcc: decode / datalist / pipeline tests, decode leg with both decoders
exec > gpurun_out/r02n_visit51.txt 2>&1
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_datalist.py tests/test_gpu_eval.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -2
run() { python bench.py --steps 3 --no-cpu-baseline --datalist-frames 0 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1: decode', round(d['e2e']['decode']['value']), 'frames/s')"; }
RPCC_BZ2_DECODER=libbz2 run "libbz2"
run "own decoder"
RPCC_BZ2_DECODER=libbz2 run "libbz2"
run "own decoder"
