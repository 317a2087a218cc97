"""Which call breaks CUDA-graph capture of the DeepFM step?  (debug helper)"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recman_b200 import _C, th
from recman_b200.th.input import DataInputs
from tests import parity_util as pu

orig_call = _C.call
def traced(name, *a):
    try:
        orig_call(name, *a)
        err = torch.cuda.is_current_stream_capturing()
    except Exception as e:
        print("C CALL FAILED:", name, e); raise
_C.call = traced
import recman_b200.ops as ops
for fuse in (False, True):
    fd = pu.make_feat_dict([50, 7, 1000, 3, 200, 31], n_dense=13)
    X, y = pu.synth_batch(fd, 256, seed=40)
    kw = dict(embedding_size=16, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=256, learning_rate=0.01,
              embedding_l2_reg=0.0, linear_l2_reg=0.0, optimizer="adagrad")
    m = th.DeepFM(fd, **kw)
    m.hparams["fuse_fm_backward"] = fuse
    first = DataInputs("cuda").load(fd, X, y)
    with torch.no_grad():
        m._out(first)
    try:
        m.compile_step(first, warmup=1)
        print("fuse", fuse, "capture ok, launches", m._graph_launches)
    except Exception as e:
        print("fuse", fuse, "capture FAILED:", str(e)[:200])
        traceback.print_exc(limit=3)
        break
