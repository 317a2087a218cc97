#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${NGPU:-2}
for ns in 0 1; do
  echo "== noscalar=$ns =="
  RM_TUNE_P2P_NOSCALAR=$ns timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/exp_$ns.err | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print(d['value'], d['ms_per_step'], {k: v['ms_per_step'] for k, v in d['kernels'].items()})
"
done
