#!/bin/bash
# One gpurun call: parity tests, smoke, short benches, ncu launch list.  Everything lands in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest ==" ; timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -x --timeout=300 > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -n 40 gpurun_out/pytest.log
echo "== smoke ==" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log ; tail -n 5 gpurun_out/smoke.log
echo "== bench small ==" ; timeout 600 python bench.py --steps 10 --warmup 3 --rows 1000000 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err ; echo "rc=$?" ; tail -n 3 gpurun_out/bench_small.err ; cat gpurun_out/bench_small.json
