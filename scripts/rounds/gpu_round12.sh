# assign with per-slice direction cones: tests, bit-for-bit against the committed library, stage times.
mkdir -p /tmp/ab
AB=r-pcc_b200/build/ab
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { local name=$1 lib=$2; shift 2; echo "== $name"; env RPCC_B200_LIB=$lib "$@" python scripts/stage_times.py 296 10 2>&1 | tail -1; env RPCC_B200_LIB=$lib "$@" python scripts/stage_times.py 1184 5 2>&1 | tail -1; env RPCC_B200_LIB=$lib "$@" python scripts/ab_ground.py dump /tmp/ab/$name.npz 2>&1 | tail -1; }
run HEAD $PWD/$AB/librpcc_HEAD.so
run cur $PWD/r-pcc_b200/lib/librpcc_b200.so
python scripts/ab_ground.py cmp /tmp/ab/HEAD.npz /tmp/ab/cur.npz | grep -c identical
