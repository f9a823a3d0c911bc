# A/B visit: FPS update batching variants + ground fit v2 against the library of HEAD (bit-for-bit + stage times).
mkdir -p /tmp/ab gpurun_out
AB=r-pcc_b200/build/ab
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
run() {  # name lib [env...]
  local name=$1 lib=$2; shift 2
  echo "== $name"
  env RPCC_B200_LIB=$lib "$@" timeout 300 python scripts/stage_times.py 296 10 2>&1 | tail -1
  env RPCC_B200_LIB=$lib "$@" timeout 300 python scripts/ab_ground.py dump /tmp/ab/$name.npz 2>&1 | tail -2
}
run HEAD $PWD/$AB/librpcc_HEAD.so
run cur $PWD/r-pcc_b200/lib/librpcc_b200.so
for v in u1 u3 u2pf u1pf; do run $v $PWD/$AB/librpcc_$v.so; done
run t512_u2 $PWD/r-pcc_b200/lib/librpcc_b200.so RPCC_FPS_THREADS=512
run t512_u4 $PWD/$AB/librpcc_u4.so RPCC_FPS_THREADS=512
run t512_u6 $PWD/$AB/librpcc_u6.so RPCC_FPS_THREADS=512
run m1_u4 $PWD/$AB/librpcc_u4.so RPCC_FPS_MINB=1
for v in cur u1 u3 u2pf u1pf t512_u2 t512_u4 t512_u6 m1_u4; do echo "-- cmp HEAD $v"; python scripts/ab_ground.py cmp /tmp/ab/HEAD.npz /tmp/ab/$v.npz; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
