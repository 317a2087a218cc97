#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for w in c5 c3 c2; do
echo "== bench $w ==" ; timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err ; echo "rc=$?"; tail -n 2 gpurun_out/bench_$w.err | cut -c1-300; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'lib', d['library_baseline'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])"
done
