#!/bin/bash
# 1-GPU call: p2p kernel tests + local-pointer microbench of the two front-end variants
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_p2p_kernels_gpu.py -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest_p2p.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest_p2p.log
tail -n 8 gpurun_out/pytest_p2p.log
python - <<'PY'
import os, torch, json
from recman_b200 import ops
B, m, k, rows = 65536, 26, 64, 4000000
table = torch.empty(m * rows, k, device="cuda").normal_(0, 0.01)
bias = torch.zeros(m * rows, device="cuda"); lin = torch.zeros(m * rows, device="cuda")
fs = torch.full((m,), rows, dtype=torch.int64, device="cuda")
lo = (torch.arange(m, dtype=torch.int64, device="cuda") * rows).contiguous()
offs = (torch.arange(m + 1, dtype=torch.int64, device="cuda") * rows).contiguous()
ids = [torch.randint(0, rows, (B, m), device="cuda") for _ in range(4)]
dense = torch.randn(B, 13, device="cuda"); ld_ = torch.randn(13, device="cuda")
def timeit(fn, n=20):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4)
res = {"gather_fm_fwd": timeit(lambda i: ops.gather_fm_fwd(table, bias, lin, offs, ids[i % 4], dense, ld_))}
for v in (0, 1):
    os.environ["RM_TUNE_P2P_ASYNC"] = str(v)
    res[f"p2p_W1_variant{v}"] = timeit(lambda i: ops.gather_fm_fwd_p2p([table.data_ptr()], [bias.data_ptr()], [lin.data_ptr()], k, fs, lo, ids[i % 4], dense, ld_))
print(json.dumps(res))
PY
