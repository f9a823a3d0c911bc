#!/bin/bash
# round 2, visit 18: the batch tool with the entropy pool choosing its coder per section (auto) vs libbz2 only
python -m pytest tests/test_gpu_datalist.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
for mode in libbz2 auto own; do
  echo "== RPCC_BZ2_CODER=$mode"; RPCC_BZ2_CODER=$mode python bench.py --steps 3 --warmup 3 --passes 1 --no-cpu-baseline 2>/dev/null | python -c 'import json,sys; b=json.loads(sys.stdin.read()); print("datalist", round(b["e2e"]["datalist"]["value"],1), "frames/s,", b["e2e"]["datalist"]["host_threads_per_rank"], "threads")'
done
