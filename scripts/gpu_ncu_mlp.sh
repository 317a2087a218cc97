#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"linear_dx_kernel|linear_dw_kernel" --launch-skip 4 -c 2 -o gpurun_out/prof_mlp -f python scripts/mlp_run.py > gpurun_out/ncu_mlp.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_mlp.log
