#!/bin/bash
# round 2, visit 22: ncu --set full of the FPS round kernel (2-CTA, records from the first-pass kernel) and of the first pass
export RPCC_FPS_NBATCH=0 RPCC_FPS_ROT=0
FRAMES=1184 REPS=4 bash scripts/ncu_full.sh r02g segment_fps_pruned_kernel fps_first_pass4_kernel
