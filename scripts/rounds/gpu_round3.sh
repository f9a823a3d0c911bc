TAG=${1:-r01g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
echo prev; RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_HEAD.so python scripts/stage_times.py
echo new; python scripts/stage_times.py
echo qocc5; RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_qocc5.so python scripts/stage_times.py
echo qocc4; RPCC_B200_LIB=$PWD/r-pcc_b200/build/ab/librpcc_qocc4.so python scripts/stage_times.py
REPS=4 bash scripts/ncu_full.sh ${TAG} ground_fit
