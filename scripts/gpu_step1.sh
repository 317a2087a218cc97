#!/bin/bash
# one 1-GPU call: GPU test-suite, kernel microbench (default + 32-byte L2 fetch granularity), C5 bench line
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --timeout=300 > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -n 15 gpurun_out/pytest.log
echo "== kbench ==" ; KB_ROWS=${KB_ROWS:-10000000} timeout 300 python scripts/kbench.py > gpurun_out/kbench.json 2> gpurun_out/kbench.err ; echo "rc=$?"; tail -n 3 gpurun_out/kbench.err; cat gpurun_out/kbench.json
echo "== kbench L2 fetch 32 ==" ; RM_TUNE_L2_FETCH=32 KB_ROWS=${KB_ROWS:-10000000} timeout 300 python scripts/kbench.py > gpurun_out/kbench_l2f32.json 2> gpurun_out/kbench_l2f32.err ; echo "rc=$?"; tail -n 3 gpurun_out/kbench_l2f32.err; cat gpurun_out/kbench_l2f32.json
echo "== bench c5 ==" ; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err ; echo "rc=$?"; tail -n 3 gpurun_out/bench_c5.err; cat gpurun_out/bench_c5.json
