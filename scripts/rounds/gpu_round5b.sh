AB=r-pcc_b200/build/ab
run() { local name=$1 lib=$2; shift 2; echo "== $name"; env RPCC_B200_LIB=$lib "$@" timeout 300 python scripts/stage_times.py 296 10 2>&1 | tail -1; }
run HEAD $PWD/$AB/librpcc_HEAD.so
run cur $PWD/r-pcc_b200/lib/librpcc_b200.so
for v in u1 u3 u2pf u1pf; do run $v $PWD/$AB/librpcc_$v.so; done
run t512_u2 $PWD/r-pcc_b200/lib/librpcc_b200.so RPCC_FPS_THREADS=512
run t512_u4 $PWD/$AB/librpcc_u4.so RPCC_FPS_THREADS=512
