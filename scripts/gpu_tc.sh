#!/bin/bash
# tcgen05 CIN bring-up: guarded by its own timeout, separate from the main suite.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== tc tests ==" ; timeout 400 python -m pytest tests/test_cin_gpu.py -m gpu -q --maxfail=60 --timeout=120 -k "tcgen05 or 3xtf32" > gpurun_out/pytest_tc.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest_tc.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/pytest_tc.log | head -60
nvidia-smi --query-gpu=name,memory.used --format=csv
for prec in ${PRECS:-3xtf32 tf32}; do
  echo "== bench c3 $prec ==" ; timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --cin-precision $prec --no-cpu-baseline > gpurun_out/bench_c3_$prec.json 2> gpurun_out/bench_c3_$prec.err ; echo "rc=$?" ; tail -n 3 gpurun_out/bench_c3_$prec.err ; cat gpurun_out/bench_c3_$prec.json
done
