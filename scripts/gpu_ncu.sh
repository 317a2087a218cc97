#!/bin/bash
# ncu evidence: (1) launch list of one eager step (every kernel + device time), (2) --set full on the two dominant kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline ${BENCH_ARGS:-}"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
RM_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
RM_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gather_fm_kernel|segment_reduce_kernel|sparse_opt" -c 4 -o gpurun_out/prof_top -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/ | head -30; tail -5 gpurun_out/ncu_full.log
