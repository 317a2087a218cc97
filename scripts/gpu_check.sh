#!/bin/bash
# One gpurun call: parity tests, smoke, short benches.  Everything lands in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest ==" ; timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --timeout=300 ${PYTEST_ARGS:-} > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -n 30 gpurun_out/pytest.log
echo "== smoke ==" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log ; tail -n 2 gpurun_out/smoke.log
for wl in ${WORKLOADS:-c5}; do
  echo "== bench $wl ${BENCH_ARGS:-} ==" ; timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err ; echo "rc=$?" ; tail -n 3 gpurun_out/bench_$wl.err ; cat gpurun_out/bench_$wl.json
done
