#!/bin/bash
# 1-GPU call: kernel tests + microbench of the segment-reduce variants
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -n 8 gpurun_out/pytest.log
echo "== kbench ==" ; KB_GATHER_VARIANTS=1 KB_ROWS=${KB_ROWS:-10000000} timeout 300 python scripts/kbench.py > gpurun_out/kbench.json 2> gpurun_out/kbench.err ; echo "rc=$?"; tail -n 3 gpurun_out/kbench.err; cat gpurun_out/kbench.json
