#!/bin/bash
# round 2, visit 47: the datalist leg (4096 files) on one box: this tree, the library before the bz2enc counter change (7104fb5),
# the library before the FPS / host-chunk changes (a050a09), coder forced to libbz2 / own, 148 frames per pipeline stage
exec > gpurun_out/r02m_visit47.txt 2>&1
run() { python bench.py --steps 3 --no-cpu-baseline --datalist-frames 4096 $2 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1: datalist', round(d['e2e']['datalist']['value']), 'frames/s  decode', round(d['e2e']['decode']['value']))"; }
run "tree"
RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_7104fb5.so run "lib 7104fb5 (before the counter change)"
RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_a050a09.so run "lib a050a09 (before wide FPS / 111-frame stages)"
RPCC_BZ2_CODER=libbz2 run "tree, libbz2 only"
RPCC_BZ2_CODER=own run "tree, own encoder only"
run "tree, host-chunk 148" "--host-chunk 148"
run "tree again"
