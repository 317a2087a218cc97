#!/bin/bash
# ncu --set full (with source) of selected kernels in one eager step.  env: KREGEX, TAG, WORKLOAD, SKIP, COUNT
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-one}
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --workload ${WORKLOAD:-c3}"
RM_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$KREGEX" --launch-skip ${SKIP:-0} -c ${COUNT:-1} -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log | cut -c1-300; ls -la gpurun_out/prof_$TAG.ncu-rep
