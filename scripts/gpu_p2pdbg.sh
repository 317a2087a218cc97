#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${NGPU:-2}
for v in 2 1 0; do
echo "== bench N=$N RM_TUNE_P2P_SCALAR=$v =="
RM_TUNE_P2P_SCALAR=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --rows 2000000 > gpurun_out/p2pdbg_$v.json 2> gpurun_out/p2pdbg_$v.err ; echo "rc=$?" ; tail -n 3 gpurun_out/p2pdbg_$v.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/p2pdbg_$v.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
PY
done
