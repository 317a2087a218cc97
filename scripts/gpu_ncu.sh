#!/bin/bash
# ncu evidence: (1) launch list of one eager step (every kernel + device time), (2) --set full on the dominant kernels.
# env: BENCH_ARGS (e.g. "--workload c3"), KREGEX (kernel-name regex for the full capture), TAG (output prefix), NFULL
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-c5}
KREGEX=${KREGEX:-"gather_fm_kernel|segment_reduce_kernel|sparse_opt"}
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline ${BENCH_ARGS:-}"
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 && \
RM_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
$CMD > gpurun_out/ncu_plain2_$TAG.log 2>&1 && \
RM_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$KREGEX" -c ${NFULL:-4} -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/ | grep $TAG; tail -5 gpurun_out/ncu_full_$TAG.log | cut -c1-300
