# plane_model_kernel v2 (cluster staged in shared memory): tests, bit-for-bit model rows against the committed library, stage times.
mkdir -p /tmp/ab
AB=r-pcc_b200/build/ab
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
RPCC_B200_LIB=$PWD/$AB/librpcc_HEAD.so AB_METHOD=plane python scripts/ab_ground.py dump /tmp/ab/HEAD.npz 2>&1 | tail -1
AB_METHOD=plane python scripts/ab_ground.py dump /tmp/ab/cur.npz 2>&1 | tail -1
python scripts/ab_ground.py cmp /tmp/ab/HEAD.npz /tmp/ab/cur.npz
python scripts/stage_times.py 296 10 uniform plane 2>&1 | tail -1
python scripts/stage_times.py 1184 5 uniform plane 2>&1 | tail -1
