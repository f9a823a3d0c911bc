# assign (squared-distance compare) + ground v2: tests, stage times (+ 512-thread ground variant), bench at N=1.
mkdir -p gpurun_out /tmp/ab
AB=r-pcc_b200/build/ab
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/stage_times.py 296 10 2>&1 | tail -1
RPCC_B200_LIB=$PWD/$AB/librpcc_gf512.so python scripts/stage_times.py 296 10 2>&1 | tail -1
RPCC_B200_LIB=$PWD/$AB/librpcc_HEAD.so python scripts/ab_ground.py dump /tmp/ab/HEAD.npz 2>&1 | tail -1
python scripts/ab_ground.py dump /tmp/ab/cur.npz 2>&1 | tail -1
RPCC_B200_LIB=$PWD/$AB/librpcc_gf512.so python scripts/ab_ground.py dump /tmp/ab/gf512.npz 2>&1 | tail -1
python scripts/ab_ground.py cmp /tmp/ab/HEAD.npz /tmp/ab/cur.npz
python scripts/ab_ground.py cmp /tmp/ab/HEAD.npz /tmp/ab/gf512.npz
timeout 600 python bench.py > gpurun_out/r01f_bench.json 2> gpurun_out/r01f_bench.err; echo "bench exit $?"; cat gpurun_out/r01f_bench.json
