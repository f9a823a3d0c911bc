# quantize with the next step's loads prefetched: slices per step / occupancy variants against the committed library.
mkdir -p /tmp/ab
AB=r-pcc_b200/build/ab
run() { local name=$1 lib=$2; echo "== $name"; RPCC_B200_LIB=$lib python scripts/stage_times.py 1184 5 2>&1 | tail -1 | sed 's/.*model=[0-9.]* //'; RPCC_B200_LIB=$lib python scripts/ab_ground.py dump /tmp/ab/$name.npz 2>&1 | tail -1; }
run HEAD $PWD/$AB/librpcc_HEAD.so
run cur $PWD/r-pcc_b200/lib/librpcc_b200.so
for v in q2 q2o5 q4o5; do run $v $PWD/$AB/librpcc_$v.so; done
for v in cur q2 q2o5 q4o5; do python scripts/ab_ground.py cmp /tmp/ab/HEAD.npz /tmp/ab/$v.npz | grep -c identical; done
