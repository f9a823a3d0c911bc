#!/bin/bash
# round 2, visit 12: stream slots of the encoder (kernels of different slots overlap at the tails of the one-CTA-per-frame launches)
for v in slots2 "" slots4 slots6; do
  if [ -n "$v" ]; then export RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_$v.so; else unset RPCC_B200_LIB; fi
  echo "== ${v:-slots3}: $(python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c 'import json,sys; b=json.loads(sys.stdin.read()); print(round(b["value"]), b["ms_per_step"])')"
done
for mb in 592 2368; do echo "== slots3 max-batch $mb: $(python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --max-batch $mb 2>/dev/null | python -c 'import json,sys; b=json.loads(sys.stdin.read()); print(round(b["value"]), b["ms_per_step"])')"; done
