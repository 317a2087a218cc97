import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recman_b200 import ops, _C, autograd as ag
from tests import parity_util as pu
from recman_b200.th import DeepFM
from recman_b200.th.input import DataInputs

def try_capture(name, fn):
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(f"{name}: capture OK", flush=True)
    except Exception as e:
        print(f"{name}: capture FAILED: {str(e).splitlines()[0]}", flush=True)
        try: torch.cuda.synchronize()
        except Exception as e2: print("sync err", e2)

class Probe(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x * 1.0
    @staticmethod
    def backward(ctx, g):
        print("  [probe bwd] capturing:", torch.cuda.is_current_stream_capturing(), "stream", torch.cuda.current_stream().cuda_stream, flush=True)
        return g

w = torch.randn(32, 32, device="cuda", requires_grad=True)
def f1():
    w.grad = None
    y = Probe.apply(w @ w).sum(); y.backward()
try_capture("probe", f1)

xbuf = torch.randn(256, 36, device="cuda", requires_grad=True)
W = torch.randn(33, 8, device="cuda", requires_grad=True); b = torch.zeros(8, device="cuda", requires_grad=True)
def f2():
    xbuf.grad = None; W.grad = None; b.grad = None
    y = ag.FirstLinearFunction.apply(xbuf, W, b, 33).sum(); y.backward()
try_capture("first_linear", f2)

fd = pu.make_feat_dict([50, 7, 100], n_dense=3)
X, y = pu.synth_batch(fd, 256, seed=1)
model = DeepFM(fd, embedding_size=16, deep_dropout=(1, 1, 1), batch_size=256, embedding_l2_reg=0.0, linear_l2_reg=0.0)
inp = DataInputs("cuda").load(fd, X, y)
def f3():
    for p in model.variables.values(): p.grad = None; p.rm_sparse_grads = []
    loss = model._loss(inp); loss.backward()
try_capture("deepfm fwd+bwd", f3)
def f4():
    model.optimizer_step()
f3()
try_capture("optimizer_step", f4)
def f5():
    model._eager_step(inp)
try_capture("full step", f5)
