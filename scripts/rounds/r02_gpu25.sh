#!/bin/bash
# round 2, visit 25: ncu --set full of the pair FPS round kernel
FRAMES=1184 REPS=4 bash scripts/ncu_full.sh r02h segment_fps_pair_kernel
