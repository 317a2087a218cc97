#!/bin/bash
# 1-GPU call: tests, the three single-GPU bench lines, ncu launch list of one eager C3 step
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -n 8 gpurun_out/pytest.log
for w in c5 c3 c2; do
echo "== bench $w ==" ; timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err ; echo "rc=$?"; tail -n 3 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json
done
TAG=c3 ; CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --workload c3"
RM_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
