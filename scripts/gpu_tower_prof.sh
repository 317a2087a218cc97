#!/bin/bash
# tower kernel micro-benchmark + one ncu --set full capture of each tower kernel at the C5 shape.  env: TAG
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-tower}
timeout 600 python scripts/tower_bench.py > gpurun_out/${TAG}_kbench.json 2> gpurun_out/${TAG}_kbench.err; echo "kbench rc=$?"; cat gpurun_out/${TAG}_kbench.json
if [ "${NCU:-1}" = "1" ]; then
KB_ONLY=fwd KB_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tower_fwd_kernel" --launch-skip 3 -c 1 -o gpurun_out/prof_${TAG}_fwd -f python scripts/tower_bench.py > gpurun_out/ncu_${TAG}_fwd.log 2>&1
echo "ncu fwd rc=$?"; tail -2 gpurun_out/ncu_${TAG}_fwd.log | cut -c1-200
KB_ONLY=bwd KB_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tower_bwd_kernel" --launch-skip 3 -c 1 -o gpurun_out/prof_${TAG}_bwd -f python scripts/tower_bench.py > gpurun_out/ncu_${TAG}_bwd.log 2>&1
echo "ncu bwd rc=$?"; tail -2 gpurun_out/ncu_${TAG}_bwd.log | cut -c1-200; ls -la gpurun_out/prof_${TAG}_*.ncu-rep
fi
