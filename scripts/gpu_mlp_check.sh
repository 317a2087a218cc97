#!/bin/bash
# 1-GPU call: full GPU suite, narrow-linear kernel timings vs cuBLAS at the C5 shape, C5 bench
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -n 8 gpurun_out/pytest.log
python - <<'PY'
import torch, json
from recman_b200 import ops
B, d, ld, N = 65536, 1677, 1680, 32
x = torch.randn(B, ld, device="cuda"); x[:, d:] = 0
W = torch.randn(d, N, device="cuda") * 0.1
g = torch.randn(B, N, device="cuda")
h = torch.randn(B, 32, device="cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4)
out = torch.empty(B, ld, device="cuda")
res = {
 "dx_ours": timeit(lambda: ops.linear_bwd_input(g, W, d_ld=ld, out=out)),
 "dx_cublas": timeit(lambda: torch.mm(g, W.t(), out=out[:, :d])),
 "dW_ours": timeit(lambda: ops.linear_bwd_weight(x, ld, d, g)),
 "dW_cublas": timeit(lambda: x[:, :d].t() @ g),
 "dW2_ours": timeit(lambda: ops.linear_bwd_weight(h, 32, 32, g)),
 "dW2_cublas": timeit(lambda: h.t() @ g),
 "dh_ours": timeit(lambda: ops.linear_bwd_input(g, W[:32].contiguous(), d_ld=32)),
 "dh_cublas": timeit(lambda: g @ W[:32].t()),
 "fwd_cublas": timeit(lambda: torch.addmm(torch.zeros(N, device="cuda"), x[:, :d], W)),
}
print(json.dumps(res))
PY
echo "== bench c5 ==" ; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err ; echo "rc=$?"; tail -n 3 gpurun_out/bench_c5.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']); [print(k, v['avg_ms'], v['ms_per_step'], v.get('gbs')) for k,v in d['kernels'].items()]"
