import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recman_b200 import ops, _C, autograd as ag
from tests import parity_util as pu
from recman_b200.th import DeepFM
from recman_b200.th.input import DataInputs

def try_capture(name, fn, mode="global"):
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(f"{name} [{mode}]: capture OK", flush=True)
    except Exception as e:
        print(f"{name} [{mode}]: capture FAILED: {str(e).splitlines()[0]}", flush=True)
        try: torch.cuda.synchronize()
        except Exception as e2: print("sync err", e2)

fd = pu.make_feat_dict([50, 7, 100], n_dense=3)
X, y = pu.synth_batch(fd, 256, seed=1)
def mk():
    model = DeepFM(fd, embedding_size=16, deep_dropout=(1, 1, 1), batch_size=256, embedding_l2_reg=0.0, linear_l2_reg=0.0)
    inp = DataInputs("cuda").load(fd, X, y)
    return model, inp

from recman_b200.autograd import pop_sparse_grads
def step_variant(model, inp, do_sparse, do_dense, do_tail):
    loss = model._loss(inp); loss.backward()
    kind, lr = 0, 0.001
    for name, p in model.variables.items():
        sparse = pop_sparse_grads(p)
        tail = getattr(p, "rm_dense_tail", None); p.rm_dense_tail = None
        if sparse:
            if do_sparse:
                for sg in sparse: ops.sparse_opt_step(p.data, sg, kind, lr, 0.0)
            if tail is not None and do_tail:
                first, g = tail
                ops.dense_opt_step(p.data.reshape(-1)[first:], g.contiguous(), kind, lr, 0.0)
        elif p.grad is not None and do_dense:
            ops.dense_opt_step(p.data, p.grad.contiguous(), kind, lr, 0.0)
        p.grad = None

for label, flags in [("sparse only", (1,0,0)), ("dense only", (0,1,0)), ("tail only", (0,0,1)), ("none", (0,0,0)), ("all", (1,1,1))]:
    model, inp = mk()
    try_capture(label, lambda: step_variant(model, inp, *flags))
model, inp = mk()
try_capture("all", lambda: step_variant(model, inp, 1, 1, 1), mode="thread_local")
model, inp = mk()
try_capture("all", lambda: step_variant(model, inp, 1, 1, 1), mode="relaxed")
