#!/bin/bash
# one call: full GPU test-suite + C3 ncu evidence
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --timeout=300 > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -n 15 gpurun_out/pytest.log
TAG=c3 BENCH_ARGS="--workload c3" KREGEX="cin_fwd_tc_kernel|cin_bwd_dx_tc_kernel|cin_bwd_dw_tc_kernel" NFULL=6 bash scripts/gpu_ncu.sh
