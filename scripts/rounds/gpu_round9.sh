# assign prefetch / occupancy variants, FPS winner fetch: stage times at 296 and 1184 frames per launch + bit-for-bit.
mkdir -p /tmp/ab
AB=r-pcc_b200/build/ab
run() { local name=$1 lib=$2; echo "== $name"; RPCC_B200_LIB=$lib python scripts/stage_times.py 296 10 2>&1 | tail -1; RPCC_B200_LIB=$lib python scripts/stage_times.py 1184 5 2>&1 | tail -1; RPCC_B200_LIB=$lib python scripts/ab_ground.py dump /tmp/ab/$name.npz 2>&1 | tail -1; }
run HEAD $PWD/$AB/librpcc_HEAD.so
run cur $PWD/r-pcc_b200/lib/librpcc_b200.so
for v in aspf0 aspf1o5 aspf0o5 fpswf0; do run $v $PWD/$AB/librpcc_$v.so; done
for v in cur aspf0 aspf1o5 aspf0o5 fpswf0; do python scripts/ab_ground.py cmp /tmp/ab/HEAD.npz /tmp/ab/$v.npz | grep -c identical; done
