#!/bin/bash
# N-GPU tuning call: p2p kernel tests, then bench with the front-end / segment-reduce variants
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m pytest tests/test_p2p_kernels_gpu.py -m gpu -q --maxfail=5 --timeout=300 2>&1 | tail -3
CFGS=${CFGS:-0:1 1:1 1:2 1:4}
for cfg in $CFGS; do
  a=${cfg%%:*}; sr=${cfg##*:}
  echo "== async=$a segred=$sr =="
  RM_TUNE_P2P_ASYNC=$a RM_TUNE_SEGRED=$sr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/tune_${a}_${sr}.err | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print(d['value'], d['ms_per_step'], {k: v['ms_per_step'] for k, v in d['kernels'].items()})
"
done
