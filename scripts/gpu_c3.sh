#!/bin/bash
# 1-GPU call: CIN tests, C3 bench line, ncu launch list of one eager C3 step
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cin_gpu.py tests/test_models_gpu.py -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest_cin.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest_cin.log
tail -n 8 gpurun_out/pytest_cin.log
echo "== bench c3 ==" ; timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err ; echo "rc=$?"; tail -n 3 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
TAG=c3 ; CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --workload c3"
RM_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
if [ "${FULL:-0}" = "1" ]; then
RM_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"cin_fwd_tc_kernel|cin_bwd_dx_tc_kernel|cin_bwd_dw_tc_kernel" -c 9 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log | cut -c1-300
fi
