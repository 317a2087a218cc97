#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi --query-gpu=index,name --format=csv
echo "== dist_check ==" ; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check.log 2>&1 ; echo "rc=$?" ; tail -n 25 gpurun_out/dist_check.log
echo "== bench N=$N ==" ; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ; echo "rc=$?" ; tail -n 5 gpurun_out/bench_n$N.err ; cat gpurun_out/bench_n$N.json
