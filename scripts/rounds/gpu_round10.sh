# FPS 64-pixel buckets (PPL=2) at 1024x2 and 512x4 against the committed library: stage times + bit-for-bit.
mkdir -p /tmp/ab
AB=r-pcc_b200/build/ab
run() { local name=$1 lib=$2; shift 2; echo "== $name"; env RPCC_B200_LIB=$lib "$@" python scripts/stage_times.py 296 10 2>&1 | tail -1; env RPCC_B200_LIB=$lib "$@" python scripts/stage_times.py 1184 5 2>&1 | tail -1; env RPCC_B200_LIB=$lib "$@" python scripts/ab_ground.py dump /tmp/ab/$name.npz 2>&1 | tail -1; }
run HEAD $PWD/$AB/librpcc_HEAD.so
run ppl1 $PWD/r-pcc_b200/lib/librpcc_b200.so
run ppl2 $PWD/r-pcc_b200/lib/librpcc_b200.so RPCC_FPS_PPL=2
run ppl2t512 $PWD/r-pcc_b200/lib/librpcc_b200.so RPCC_FPS_PPL=2 RPCC_FPS_THREADS=512
for v in ppl1 ppl2 ppl2t512; do python scripts/ab_ground.py cmp /tmp/ab/HEAD.npz /tmp/ab/$v.npz | grep -c identical; done
