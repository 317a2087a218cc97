#!/bin/bash
# Round-2 evidence (one GPU): ncu launch lists of one eager C5 / C3 / C2 bench run and one `ncu --set full` capture of
# the tower / head kernels (C5) - each only after the same command has exited 0 without ncu.  Outputs -> gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
FLAGS="--steps 2 --warmup 3 --blocks 1 --no-graph --no-other-workloads --no-cpu-baseline --no-library-baseline"
for W in c5 c3 c2; do
  timeout 600 python bench.py --workload $W $FLAGS > gpurun_out/r2_plain_$W.json 2> gpurun_out/r2_plain_$W.err; rc=$?
  echo "plain $W rc=$rc"
  if [ $rc -eq 0 ]; then
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_$W.csv \
      python bench.py --workload $W $FLAGS > gpurun_out/r2_ncu_launches_$W.log 2>&1
    echo "ncu launches $W rc=$? lines=$(wc -l < gpurun_out/r2_launches_$W.csv)"
  fi
done
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"tower_fwd_kernel|tower_bwd_kernel|head_kernel" --launch-skip 9 -c 3 -o gpurun_out/prof_r2_final_c5 -f \
  python bench.py --workload c5 $FLAGS > gpurun_out/r2_ncu_full_c5.log 2>&1
echo "ncu full c5 rc=$?"; tail -2 gpurun_out/r2_ncu_full_c5.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"cin_fwd_tc_kernel|cin_bwd_dx_tc_kernel|cin_bwd_dw_tc_kernel|cin_splitk_finish" --launch-skip 20 -c 6 -o gpurun_out/prof_r2_final_c3 -f \
  python bench.py --workload c3 $FLAGS > gpurun_out/r2_ncu_full_c3.log 2>&1
echo "ncu full c3 rc=$?"; tail -2 gpurun_out/r2_ncu_full_c3.log | cut -c1-200
ls -la gpurun_out/prof_r2_final_*.ncu-rep
