#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi --query-gpu=index,name --format=csv
nvidia-smi topo -m 2>/dev/null | head -12
for mode in ${MODES:-p2p a2a}; do
  echo "== dist_check $mode ==" ; DIST_MODE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check_$mode.log 2>&1 ; echo "rc=$?" ; tail -n 12 gpurun_out/dist_check_$mode.log
done
echo "== bench N=$N ==" ; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ; echo "rc=$?" ; tail -n 5 gpurun_out/bench_n$N.err ; cat gpurun_out/bench_n$N.json
if [ "${ALSO_EAGER:-1}" = "1" ]; then
echo "== bench N=$N eager ==" ; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-graph ${BENCH_ARGS:-} > gpurun_out/bench_n${N}_eager.json 2> gpurun_out/bench_n${N}_eager.err ; echo "rc=$?" ; tail -n 5 gpurun_out/bench_n${N}_eager.err ; cat gpurun_out/bench_n${N}_eager.json
fi
