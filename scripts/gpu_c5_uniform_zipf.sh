#!/bin/bash
# 1-GPU call: full GPU suite, C5 bench with uniform and Zipf ids
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -n 12 gpurun_out/pytest.log
for ids in uniform zipf; do
echo "== bench c5 $ids ==" ; timeout 600 python bench.py --ids $ids --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c5_$ids.json 2> gpurun_out/bench_c5_$ids.err ; echo "rc=$?"; tail -n 2 gpurun_out/bench_c5_$ids.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_c5_$ids.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value']); [print('  ', k, v['avg_ms'], v['ms_per_step'], v.get('gbs')) for k,v in d['kernels'].items()]"
done
