#!/usr/bin/env python
"""bench.py - CTR train samples/sec on synthetic Criteo-shaped data (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: encode -> fused gather+FM+first-order front end ->
MLP (cuBLAS) -> loss -> backward -> deterministic sparse scatter-add -> fresh-optimizer update.

Default workload (config.workload) = C5: DeepFM, 26 sparse fields x 10 M rows x k=64 fp32 tables (the C4 tables,
66.6 GB, row-sharded over the ranks), 13 dense, batch 65 536 per GPU (weak scaling), deep tower (32, 32) (the
reference's DeepFM default), uniform ids.  Other workloads: c2 = DCN 6 cross + 3x400 MLP k=16 B=4096,
c3 = xDeepFM CIN 200-200-200 k=16 B=8192.

One JSON line on stdout (rank 0).  `value` = device-timed whole-job samples/s with inputs resident in HBM;
`e2e` = the same through the public call with pinned HOST buffers (H2D + D2H loss inside the timed region).
`--impl reference` times the CPU oracle (the only runnable "reference CPU path", see oracle/__init__.py).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # name: model, fields, rows/table, k, n_dense, per-GPU batch, extra
    "c5": dict(model="DeepFM", m=26, rows=10_000_000, k=64, n_dense=13, batch=65536, deep=(32, 32),
               desc="C5 DeepFM fwd+bwd+opt, 26 x 10M x k=64 fp32 tables (C4), B=65536/GPU, deterministic scatter-add"),
    "c2": dict(model="DCN", m=26, rows=1_000_000, k=16, n_dense=13, batch=4096, deep=(400, 400, 400), cross=6,
               desc="C2 DCN 6 cross + 3x400 MLP, 26 x 1M x k=16, B=4096"),
    "c3": dict(model="xDeepFM", m=26, rows=1_000_000, k=16, n_dense=13, batch=8192, deep=(400, 400),
               cin=(200, 200, 200), desc="C3 xDeepFM CIN 200-200-200 + DNN, 26 x 1M x k=16, B=8192"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=None, help="rows per table (override)")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (override)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): the workload's batch per GPU; strong: the workload's batch is the GLOBAL batch")
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--ids", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the torch-eager-CUDA comparison")
    ap.add_argument("--no-graph", action="store_true", help="do not CUDA-graph the step")
    ap.add_argument("--cin-precision", default="3xtf32")
    ap.add_argument("--blocks", type=int, default=5, help="timed blocks of --steps steps each; the median block is reported")
    ap.add_argument("--no-other-workloads", action="store_true", help="c5 only: skip the c2 / c3 lines")
    ap.add_argument("--cpu-batch", type=int, default=8192)
    ap.add_argument("--cpu-rows", type=int, default=100_000)
    return ap.parse_args()


# --------------------------------------------------------------------------------------- data
def make_ids(rng: np.random.RandomState, B, m, rows, dist):
    if dist == "uniform":
        return rng.randint(0, rows, size=(B, m)).astype(np.int64)
    z = rng.zipf(1.05, size=(B, m)).astype(np.int64) - 1
    return z % rows


def make_batch(seed, B, m, rows, n_dense, dist):
    rng = np.random.RandomState(seed)
    ids = make_ids(rng, B, m, rows, dist)
    dense = rng.randn(B, n_dense).astype(np.float32)
    y = (rng.rand(B) < 0.25).astype(np.float32)
    return ids, dense, y


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- our arm
def build_model(w, args, world, rank):
    from recman_b200.th import DCN, DeepFM, xDeepFM
    from recman_b200.th.input import DenseFeat, FeatureDictionary, SparseFeat

    rows = args.rows or w["rows"]
    k = args.k or w["k"]
    fd = FeatureDictionary()
    for i in range(w["m"]):
        fd[f"C{i}"] = SparseFeat(f"C{i}", rows - 1, encoder=False)
    for j in range(w["n_dense"]):
        fd[f"I{j}"] = DenseFeat(f"I{j}", scaler=False)
    B = args.batch or w["batch"]
    if args.scaling == "strong":  # fixed global batch, split evenly over the ranks
        B = max(1, B // world)
    common = dict(embedding_size=k, embedding_l2_reg=0.0, linear_l2_reg=0.0, batch_size=B, learning_rate=1e-3,
                  optimizer="adam")
    if w["model"] == "DeepFM":
        model = DeepFM(fd, deep_hidden_units=w["deep"], deep_dropout=(1.0,) * (len(w["deep"]) + 1), deep_l2_reg=0.0,
                       **common)
    elif w["model"] == "DCN":
        model = DCN(fd, deep_hidden_units=w["deep"], deep_dropout=(1.0,) * (len(w["deep"]) + 1),
                    cross_layer_num=w["cross"], **common)
    else:
        from recman_b200.th.layers import leaky_relu

        hp = dict(embedding_size=k, embedding_l2_reg=0.0, linear_l2_reg=0.0, deep_hidden_units=w["deep"],
                  deep_dropout=(1.0,) * (len(w["deep"]) + 1), deep_l2_reg=0.0, cin_cross_layer_units=list(w["cin"]),
                  cin_dropout=[1] * (len(w["cin"]) + 1), cin_l2_reg=0.0, learning_rate=1e-3, optimizer="adam",
                  cin_precision=args.cin_precision, deep_activation=leaky_relu, cin_activation=leaky_relu)
        model = xDeepFM(fd, hp, batch_size=B)
    return model, fd, rows, k, B


def algorithmic_bytes(kind, B, m, k, n_dense, n_unique=None):
    """Algorithmic HBM bytes per launch."""
    if kind in ("rm_gather_fm_fwd", "rm_gather_fm_fwd_p2p"):
        # ids 8 + row read 4k + row write 4k + bias 4 + lin 4 per (b,f); dense read+write; S write; 2 logits
        return B * (m * (8 + 8 * k + 8) + 8 * n_dense + 4 * k + 8)
    if kind == "rm_emb_fm_bwd":
        # per id: position 4 + dx row 4k + x row 4k; S row + 2 scalars per sample; per unique row: 4k + 2*4 written
        nu = n_unique if n_unique is not None else B * m
        return B * m * (4 + 8 * k) + B * (4 * k + 8) + nu * (4 * k + 8)
    if kind == "rm_emb_fm_bwd_update":
        # rm_emb_fm_bwd with the optimizer fused in: instead of writing the summed rows, per unique row the row id (8) is
        # read and the table row (4k) and its bias / linear weights (2*4) are read AND written
        nu = n_unique if n_unique is not None else B * m
        return B * m * (4 + 8 * k) + B * (4 * k + 8) + nu * (8 + 2 * (4 * k + 8))
    if kind == "rm_segment_reduce_p2p_update":
        # per owned id: global position 4 + gradient row (k+4)*4 (local or NVLink); per unique row as above
        nu = n_unique if n_unique is not None else B * m
        return B * m * (4 + 4 * (k + 4)) + nu * (8 + 2 * (4 * k + 8))
    if kind == "rm_segment_reduce_p2p":
        # per owned id: global position 4 + gradient row (k+4)*4 (local or NVLink); per unique row: 4k + 2*4 written
        nu = n_unique if n_unique is not None else B * m
        return B * m * (4 + 4 * (k + 4)) + nu * (4 * k + 8)
    if kind == "rm_pack_grad_rows":
        # per (b,f): dx row + x row read, G row (k+4) written; S row + 2 scalars per sample
        return B * m * (8 * k + 4 * (k + 4)) + B * (4 * k + 8)
    if kind in ("rm_tower_fwd", "rm_tower_fwd_p2p"):
        # per (b,f): id 8 + row read 4k + interleaved (bias, weight) 8; per sample: dense read, S + y1 + 2 logits written
        # (N1 = 32 at the bench shape; the row buffer is not written)
        return B * (m * (8 + 4 * k + 8) + 4 * n_dense + 4 * k + 4 * 32 + 8)
    if kind == "rm_tower_bwd_update":
        # per position: sorted key + position 8, table row read 4k; per sample (read once): g1 row 4*32, S row 4k,
        # g_fm + g_lin 8; per unique row: table row written 4k, (bias, weight) read + written 16
        nu = n_unique if n_unique is not None else B * m
        return B * m * (8 + 4 * k) + B * (4 * 32 + 4 * k + 8) + nu * (4 * k + 16)
    if kind == "rm_cross_fwd":
        # the fused cross network reads the row and writes nothing but L dots + the logit: 2*4*d per sample
        return B * 8 * (m * k + n_dense)
    if kind == "rm_cross_bwd":
        return B * 12 * (m * k + n_dense)  # x0 read, upstream gradient read, dx written: 3*4*d per sample
    # rm_sparse_opt_step is launched once per table kind (k = 64, 1, 1): its averaged time is not a single-kernel
    # figure, so it is listed without bytes
    return None


def cin_flops_per_step(w, B, k):
    """Algorithmic flops of the CIN contraction per step: fwd 2*B*D*m*H_i*N_i per layer; bwd = 2x fwd."""
    if w.get("cin") is None:
        return None
    m, D = w["m"], k
    H, total = m, 0
    units = list(w["cin"])
    for i, N in enumerate(units):
        total += 2 * B * D * m * H * N
        H = N // 2 if i < len(units) - 1 else N  # split-half: first half feeds the next layer
    return total


def measure_tf32_peak(dev, n=8192, iters=20):
    """Dense TF32 tensor-core throughput of this GPU, measured live: cuBLAS TF32 GEMM n^3 (a library call used ONLY as
    the denominator of the CIN kernels' issued-MMA fraction; BASELINE.md section 2 promised a measured figure)."""
    a = torch.randn(n, n, device=dev)
    b = torch.randn(n, n, device=dev)
    c = torch.empty(n, n, device=dev)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(3):
            torch.mm(a, b, out=c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            torch.mm(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    del a, b, c
    return 2.0 * n ** 3 * iters / (ms * 1e-3) / 1e12


def measure(args, wname, world, rank, local_rank, dev, primary=True):
    """Build the workload's model, time it (device-resident and end-to-end) and return the JSON-able record.
    `primary` adds the CPU / library baselines; secondary workloads (other_workloads) are timed more briefly."""
    import torch.distributed as dist

    w = WORKLOADS[wname]
    from recman_b200 import ops
    from recman_b200.th.input import DataInputs, HostPrefetcher

    model, fd, rows, k, B = build_model(w, args, world, rank)
    if world > 1:
        from recman_b200.th import dist as rdist

        rdist.shard_model(model, world, rank)
    m, n_dense = w["m"], w["n_dense"]
    steps = args.steps if primary else max(10, min(args.steps, 20))
    blocks = max(1, args.blocks if primary else min(args.blocks, 3))

    # rotating pool of distinct batches (ids differ per step so nothing is served from L2 by repetition)
    NB = 8
    host = [make_batch(2019 + 1000 * rank + i, B, m, rows, n_dense, args.ids) for i in range(NB)]
    pinned = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory(), torch.from_numpy(c).pin_memory())
              for a, b, c in host]
    resident = [DataInputs.from_tensors(fd, a.to(dev), b.to(dev), c.to(dev)) for a, b, c in pinned]

    def step_resident(i):
        return model.fit_on_batch(resident[i % NB], None)

    # e2e input pipeline: every step's batch is copied from pinned host memory (double-buffered: the copy of step i+1
    # rides under step i on a copy stream); the loss is read back to the host every step
    prefetch = HostPrefetcher(fd, lambda i: pinned[i % NB], dev)

    # every step's loss is copied to pinned host memory and read by the host - one step late, so that the host can
    # enqueue step i+1 while step i runs (the timed region ends with a full synchronize: the last loss has landed too)
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [None, None]
    e2e_losses = []

    def step_e2e(i):
        loss = model.fit_on_batch(prefetch.get(i), None)
        slot = i & 1
        loss_host[slot].copy_(loss.detach(), non_blocking=True)  # D2H read of the step's result
        loss_ev[slot] = torch.cuda.Event()
        loss_ev[slot].record()
        if loss_ev[slot ^ 1] is not None:
            loss_ev[slot ^ 1].synchronize()
            e2e_losses.append(float(loss_host[slot ^ 1]))

    # ---- warm-up (also creates the variables) ----
    for i in range(args.warmup):
        step_resident(i)
    torch.cuda.synchronize()
    model.check_ids()

    # ---- per-kernel CUDA-event profile of the same steps, eager (kernel durations do not depend on how they
    #      are launched); the timed region below replays the CUDA graph of the step ----
    ops.enable_profile()
    launches0 = ops.launch_count()
    prof_steps = min(steps, 10)
    for i in range(prof_steps):
        step_resident(i)
    launches_per_step = (ops.launch_count() - launches0) // prof_steps
    prof = ops.disable_profile()
    graphed = False
    if not args.no_graph and (world == 1 or model.shard.peer is not None):
        model.compile_step(resident[0], warmup=1)  # peer-memory sharding has no host sync: capturable with NCCL inside
        graphed = True
        for i in range(args.warmup):
            step_resident(i)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(nsteps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ncu_range = os.environ.get("RM_NCU_RANGE") == "1"  # ncu --profile-from-start off: profile the timed region only
    if ncu_range:
        torch.cuda.profiler.start()
    # `blocks` timed blocks of exactly `steps` steps each (barrier + synchronize on both sides, max over ranks);
    # the reported step time is the median block
    block_ms, block_wall = [], []
    for _ in range(blocks):
        ms_b, wall_b = timed(step_resident, steps)
        block_ms.append(ms_b)
        block_wall.append(wall_b)
    if ncu_range:
        torch.cuda.profiler.stop()
    launches = launches_per_step * steps
    ms_total = statistics.median(block_ms)
    wall_total = statistics.median(block_wall)
    ms_step = ms_total / steps
    value = B * world / (ms_step * 1e-3)

    # ---- e2e: pinned host buffers -> H2D -> step -> D2H loss ----
    for i in range(2):
        step_e2e(i)
    e2e_blocks = []
    for _ in range(blocks):
        prefetch._pending = None  # every block starts cold: its first batch is copied inside it
        loss_ev[0] = loss_ev[1] = None
        e2e_ms, e2e_wall = timed(step_e2e, steps)
        e2e_blocks.append(max(e2e_ms, e2e_wall))  # includes the host wait for the D2H read
    clocks = sampler.stop() if rank == 0 else None
    assert all(np.isfinite(v) for v in e2e_losses)
    e2e_ms_step = statistics.median(e2e_blocks) / steps
    e2e_value = B * world / (e2e_ms_step * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in pinned[0])

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        n_unique = int(np.unique(host[0][0] + (np.arange(m, dtype=np.int64) * rows)[None, :]).size)
        kernels = {}
        for name, (total_ms, count) in prof.items():
            avg = total_ms / max(count, 1)
            ab = algorithmic_bytes(name, B, m, k, n_dense, n_unique)
            kernels[name] = {"avg_ms": round(avg, 4), "launches": count, "ms_per_step": round(total_ms / prof_steps, 4),
                             "alg_bytes": ab, "gbs": (round(ab / (avg * 1e-3) / 1e9, 1) if ab and avg > 0 else None)}
            if world > 1 and name in ("rm_tower_fwd_p2p", "rm_gather_fm_fwd_p2p") and avg > 0:
                # rows (and their (bias, weight) pairs) read from the W-1 peers over NVLink, per launch (uniform ids)
                nv = B * m * (world - 1) / world * (4 * k + 8)
                kernels[name]["nvlink_read_bytes"] = int(nv)
                kernels[name]["nvlink_gbs"] = round(nv / (avg * 1e-3) / 1e9, 1)
        if world > 1 and "rm_tower_bwd_update" in kernels:
            # what the collectives of the step deliver to this rank: ids (int32) + per-sample operands (S, g1, g_fm, g_lin)
            kernels["nccl_all_gather"] = {"avg_ms": None, "ms_per_step": None, "alg_bytes": None, "recv_bytes_per_step": int((world - 1) * B * (4 * m + 4 * k + 4 * 32 + 8)),
                                          "note": "ids on the side stream under the forward; operands before the backward"}
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wname, {})
        except Exception:
            pass
        cand = [(v["ms_per_step"], n) for n, v in kernels.items() if v.get("alg_bytes")]
        roofline = None
        cin_f = cin_flops_per_step(w, B, k)
        if cin_f and "rm_cin_layer_bwd" in kernels:
            # the CIN contraction dominates xDeepFM: tensor-pipe roofline.  Peak = measured dense bf16 (the number
            # MEASURED_PEAKS.json holds); kind::tf32 runs at half the bf16 rate and the 3xTF32 parity mode issues 3 MMAs
            # per algorithmic product, so frac understates pipe occupancy (ncu: profiles/*c3*_ncu_full.json).
            tpeak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1376.3)))  # timed inside a step
            fwd_ms, bwd_ms = kernels["rm_cin_layer_fwd"]["ms_per_step"], kernels["rm_cin_layer_bwd"]["ms_per_step"]
            for nm, fl, ms in (("rm_cin_layer_fwd", cin_f, fwd_ms), ("rm_cin_layer_bwd", 2 * cin_f, bwd_ms)):
                kernels[nm]["alg_flops_per_step"] = fl
                kernels[nm]["tflops"] = round(fl / (ms * 1e-3) / 1e12, 1) if ms > 0 else None
            passes = 3 if args.cin_precision == "3xtf32" else 1
            dom = "rm_cin_layer_bwd" if bwd_ms >= fwd_ms else "rm_cin_layer_fwd"
            ach = kernels[dom]["tflops"]
            try:
                tf32_peak, tf32_src = measure_tf32_peak(dev), "measured in this run: cuBLAS TF32 GEMM 8192^3 x 20"
            except Exception as exc:  # the denominator only: fall back to half the bf16 figure
                tf32_peak, tf32_src = tpeak / 2, f"assumed = bf16 / 2 ({type(exc).__name__})"
            roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s",
                        "frac": round(ach / tpeak, 4), "peak_source": "measured dense bf16, sustained (MEASURED_PEAKS.json)"
                        if peaks else "fallback 1376 TFLOP/s", "traffic": (traffic.get(dom) or {}).get("bytes"),
                        "traffic_source": (traffic.get(dom) or {}).get("source"),
                        "mma_passes": passes, "issued_tflops": round(ach * passes, 1),
                        "tf32_peak_tflops": round(tf32_peak, 1), "tf32_peak_source": tf32_src,
                        "frac_of_tf32_peak_issued": round(ach * passes / tf32_peak, 4),
                        "avg_launch_ms": kernels[dom]["avg_ms"], "alg_flops_per_step": kernels[dom]["alg_flops_per_step"]}
        elif cand:
            _, dom = max(cand)
            kv = kernels[dom]
            roofline = {"kernel": dom, "bound": "hbm", "achieved": kv["gbs"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(kv["gbs"] / hbm_peak, 4), "frac_of_nominal_8000": round(kv["gbs"] / 8000.0, 4),
                        "peak_source": peak_src, "traffic": (traffic.get(dom) or {}).get("bytes"),
                        "traffic_source": (traffic.get(dom) or {}).get("source"), "avg_launch_ms": kv["avg_ms"],
                        "alg_bytes_per_launch": kv["alg_bytes"]}
        cpu = None
        if primary and not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline(args, w, steps=3, warmup=1)
        lib = None
        if not args.no_library_baseline and world == 1:
            try:
                lib = library_baseline(args, w, model, resident, B, m, k, rows, n_dense)
            except Exception as exc:  # a reported comparison, never a reason to lose the bench line
                lib = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        out = {
            "metric": "CTR train samples/sec", "value": round(value, 1), "unit": "samples/s", "n_gpus": world,
            "steps": steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "model": w["model"], "fields": m, "rows_per_table": rows, "k": k,
                       "n_dense": n_dense, "batch_per_gpu": B, "global_batch": B * world, "ids": args.ids,
                       "optimizer": "adam (fresh per batch, as the reference)", "l2_flush":
                       "inputs larger than L2: tables %.1f GB, 8 rotating id batches" % (m * rows * k * 4 / 1e9),
                       "parallelism": ("single GPU" if world == 1 else f"row-sharded tables x{world} "
                                       f"({'NVLink peer-memory row reads, owner-side update' if model.shard.peer is not None else 'NCCL all-to-all'})"
                                       " + DP dense all-reduce"),
                       "step_launch": ("CUDA graph replay" if graphed else "eager"),
                       "timing": f"median of {blocks} blocks of {steps} steps; each block bracketed by barrier + "
                                 "synchronize, CUDA events, max over ranks",
                       "front_end": ("fused tower: rm_tower_fwd / rm_tower_bwd_update (tcgen05)" if "rm_tower_fwd" in kernels
                                     else "fused tower over peer memory: rm_tower_fwd_p2p / rm_tower_shard_plan / "
                                          "rm_tower_bwd_update (tcgen05)" if "rm_tower_fwd_p2p" in kernels
                                     else "rm_gather_fm_fwd + separate backward kernels")},
            "clocks": clocks,
            "block_ms": [round(v, 3) for v in block_ms],
            "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "ms_per_step": round(e2e_ms_step, 4),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "input_pipeline": "pinned host buffers, double-buffered H2D on a copy stream (HostPrefetcher); "
                                      "every step's loss copied to pinned host memory and read by the host one step late"},
            "gpu_launches": launches,
            "roofline": roofline,
            "kernels": kernels,
            "cpu_baseline": cpu,
            "library_baseline": lib,
            "wall_ms_per_step": round(wall_total / steps, 4),
        }
    # release the workload's device memory before the next one is built
    del model, resident, prefetch, pinned, host
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    parity = None
    if world > 1:
        try:
            parity = sharded_parity_check(world, rank, dev)
        except Exception as exc:
            parity = f"FAILED: {type(exc).__name__}: {exc}"[:400]
    out = measure(args, args.workload, world, rank, local_rank, dev, primary=True)
    if rank == 0 and parity is not None:
        out["parity_check"] = parity
    if world == 1 and args.workload == "c5" and not args.no_other_workloads and not (args.rows or args.batch or args.k):
        # BASELINE.json configs[1] and configs[2] in the same run, timed more briefly (their own clocks / roofline)
        others = {}
        for name in ("c2", "c3"):
            try:
                others[name] = measure(args, name, world, rank, local_rank, dev, primary=False)
            except Exception as exc:
                others[name] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        out["other_workloads"] = others
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        # a captured graph that holds NCCL kernels plus cudaIpc peer mappings makes the orderly teardown
        # (destroy_process_group / graph destruction) wait on the peers: leave without it
        os._exit(0)
    return out


def sharded_parity_check(world, rank, dev, k=64, b=256):
    """Before timing at N > 1: one batch through a row-sharded DeepFM (small tables, same kernels and exchange as the
    timed model) against the single-GPU model on the concatenated global batch - logits per sample, and the
    parameters after one optimizer step.  Product code on both sides (no oracle).  Returns "ok" or raises."""
    import torch.distributed as dist

    from recman_b200.th import DeepFM
    from recman_b200.th import dist as rdist
    from recman_b200.th.input import DataInputs, DenseFeat, FeatureDictionary, SparseFeat

    sizes = [50, 7, 1000, 3, 200, 31, 2, 90, 1]
    n_dense = 5
    fd = FeatureDictionary()
    for i, v in enumerate(sizes):
        fd[f"C{i}"] = SparseFeat(f"C{i}", v - 1, encoder=False)
    for j in range(n_dense):
        fd[f"I{j}"] = DenseFeat(f"I{j}", scaler=False)
    kw = dict(embedding_size=k, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=b, embedding_l2_reg=0.0,
              linear_l2_reg=0.0, deep_l2_reg=0.0, learning_rate=0.05, optimizer="gd")

    def batch(r):
        rng = np.random.RandomState(100 + r)
        ids = np.stack([rng.randint(0, v, size=b) for v in sizes], 1).astype(np.int64)
        return ids, rng.randn(b, n_dense).astype(np.float32), (rng.rand(b) < 0.3).astype(np.float32)

    parts = [batch(r) for r in range(world)]
    glob = tuple(np.concatenate([p[i] for p in parts]) for i in range(3))
    to_inputs = lambda t: DataInputs.from_tensors(fd, torch.from_numpy(t[0]).to(dev), torch.from_numpy(t[1]).to(dev),
                                                  torch.from_numpy(t[2]).to(dev))
    ref = DeepFM(fd, **dict(kw, batch_size=b * world))
    ref.hparams["tower"] = False  # single GPU: the separate kernels; sharded: the fused tower (k = 64) over peer memory
    gi = to_inputs(glob)
    with torch.no_grad():
        ref._out(gi)
    g = torch.Generator().manual_seed(1)
    for name, p in ref.variables.items():
        p.data.copy_((torch.randn(p.shape, generator=g) * 0.05).to(dev))
    model = DeepFM(fd, **kw)
    rdist.shard_model(model, world, rank)
    li = to_inputs(parts[rank])
    with torch.no_grad():
        model._out(li)
    plan = model.shard
    offs = np.concatenate([[0], np.cumsum(sizes)])
    total = int(offs[-1])
    rows_of = [plan.local_rows_of(f) + int(offs[f]) for f in range(len(sizes))]

    for name, p in model.variables.items():
        gten = ref.variables[name].data
        if name in ("feat_embed_table", "feat_bias_table", "linear_w"):
            for f, rows in enumerate(rows_of):
                lo = plan.local_offsets[f]
                p.data[lo : lo + rows.numel()] = gten[rows.to(dev)]
            if name == "linear_w":
                p.data[plan.total_local :] = gten[total:]
        else:
            p.data.copy_(gten)
    # the no_grad forward below reads the peers' shards and has no collective in front of it: every rank must have
    # filled its shard first (inside a training step the ids all-gather is that barrier)
    torch.cuda.synchronize()
    dist.barrier()
    errs = []  # collected, not raised: a rank that leaves early would strand the others in the step's collectives

    def check(what, got, exp, atol):
        try:
            torch.testing.assert_close(got, exp, rtol=1e-5, atol=atol)
        except AssertionError as exc:
            errs.append(f"{what}: " + " ".join(str(exc).split())[:200])

    with torch.no_grad():
        lg = ref._out(gi, training=True)
        ll = model._out(li, training=True)
    check("logits", ll, lg[rank * b : (rank + 1) * b], 2e-6)
    ref.fit_on_batch(gi, None)
    model.fit_on_batch(li, None)
    torch.cuda.synchronize()
    model.check_ids()
    for name, p in model.variables.items():
        gten = ref.variables[name].data
        if name in ("feat_embed_table", "feat_bias_table", "linear_w"):
            for f, rows in enumerate(rows_of):
                lo = plan.local_offsets[f]
                exp = gten[rows.to(dev)]
                check(f"{name}[table {f}]", p.data[lo : lo + rows.numel()], exp,
                      1e-5 * max(float(exp.abs().max()) if exp.numel() else 0.0, 1e-3))
        else:
            check(name, p.data, gten, 1e-5 * max(float(gten.abs().max()), 1e-3))
    bad = torch.tensor([len(errs)], dtype=torch.int64, device=dev)
    dist.all_reduce(bad)
    del model, ref
    torch.cuda.empty_cache()
    if int(bad.item()):  # same verdict on every rank
        raise AssertionError(f"sharded parity: {int(bad.item())} mismatches over all ranks; rank {rank}: {errs[:3]}")
    return f"ok: row-sharded DeepFM x{world} (k={k}, b={b}/rank) == single-GPU on the global batch: logits 1e-5, all parameters after one step 1e-5"


# --------------------------------------------------------------------------------------- library bar (torch eager)
def library_baseline(args, w, model, resident, B, m, k, rows, n_dense, steps=5, warmup=2):
    """the same step written with stock PyTorch CUDA ops (F.embedding with sparse gradients, cuBLAS,
    autograd, coalesce + index updates) on the same B200 and the SAME tables - "library kernels", the bar the
    hand-written path has to beat.  Not the oracle (nothing from oracle/ is used) and not part of the product."""
    import torch.nn.functional as F

    dev = resident[0].sparse_ids.device
    v = model.variables
    T = v["feat_embed_table"].data.requires_grad_()
    total = m * rows
    # own contiguous copies of the k=1 tables (the fused tower keeps them interleaved in one [rows, 2] array)
    Bt = v["feat_bias_table"].data.reshape(-1, 1).clone().requires_grad_() if "feat_bias_table" in v else None
    lw = v["linear_w"].data
    Lt = lw[:total].reshape(-1, 1).clone().requires_grad_()
    offs = (torch.arange(m, device=dev, dtype=torch.int64) * rows)[None, :]
    g = torch.Generator(device=dev).manual_seed(7)
    d = m * k + n_dense
    dims = [d] + list(w["deep"])
    dense_p = []

    def P(*shape, scale=0.05):
        t = (torch.randn(*shape, device=dev, generator=g) * scale).requires_grad_()
        dense_p.append(t)
        return t

    Ws = [P(dims[i], dims[i + 1]) for i in range(len(w["deep"]))]
    bs = [P(dims[i + 1], scale=0.0) for i in range(len(w["deep"]))]
    w_out, w0, lin_dense = P(dims[-1], 1), P(1, scale=0.0), P(max(n_dense, 1), scale=0.0)
    act = torch.relu if w["model"] != "xDeepFM" else (lambda t: F.leaky_relu(t, 0.2))
    if w["model"] == "DCN":
        L = w["cross"]
        cw, cb, co = P(L, d), P(L, d, scale=0.0), P(d, 1)
    if w["model"] == "xDeepFM":
        filt, fb, H, tot = [], [], m, 0
        units = list(w["cin"])
        for i, N in enumerate(units):
            filt.append(P(m * H, N))
            fb.append(P(N, scale=0.0))
            keep = N if i == len(units) - 1 else N - N // 2
            tot += keep
            H = N // 2 if i < len(units) - 1 else N
        cin_w = P(tot, 1)
    lr = 1e-3
    lr_t = lr * (1 - 0.999) ** 0.5 / (1 - 0.9)

    def adam1(p, gr):  # fresh Adam, first step (the reference creates a new optimizer every batch)
        return p - lr_t * (0.1 * gr) / ((0.001 * gr * gr).sqrt() + 1e-7)

    def step(i):
        inp = resident[i % len(resident)]
        keys = inp.sparse_ids + offs
        e = F.embedding(keys, T, sparse=True)
        lin = F.embedding(keys, Lt, sparse=True).sum((1, 2))
        dn = inp.dense
        if n_dense:
            lin = lin + dn @ lin_dense[:n_dense]
        x = torch.cat([e.flatten(1), dn], 1) if n_dense else e.flatten(1)
        h = x
        for W_, b_ in zip(Ws, bs):
            h = act(torch.addmm(b_, h, W_))
        logit = lin + (h @ w_out).squeeze(1) + w0
        if w["model"] == "DeepFM":
            S = e.sum(1)
            logit = logit + F.embedding(keys, Bt, sparse=True).sum((1, 2)) + 0.5 * ((S * S).sum(1) - (e * e).sum((1, 2)))
        elif w["model"] == "DCN":
            xl = x
            for l in range(w["cross"]):
                xl = x * (xl @ cw[l]).unsqueeze(1) + cb[l] + xl
            logit = logit + (xl @ co).squeeze(1)
        else:
            x0, xk, outs = e, e, []
            for i_, (Wf, bf) in enumerate(zip(filt, fb)):
                z = torch.einsum("bpd,bqd->bdpq", x0, xk).flatten(2)           # [B, D, m*H]
                f = act(z @ Wf + bf).transpose(1, 2)                            # [B, N, D]
                N = f.shape[1]
                if i_ < len(filt) - 1:
                    xk, keep = f[:, : N // 2], f[:, N // 2 :]
                else:
                    keep = f
                outs.append(keep.sum(2))
            logit = logit + (torch.cat(outs, 1) @ cin_w).squeeze(1)
        pred = torch.sigmoid(logit).clamp(1e-7, 1 - 1e-7)
        y = inp["y"]
        loss = -(y * torch.log(pred + 1e-7) + (1 - y) * torch.log(1 - pred + 1e-7)).mean()
        loss.backward()
        with torch.no_grad():
            for t in (T, Bt, Lt):
                if t is None or t.grad is None:
                    continue
                gsp = t.grad.coalesce()
                idx = gsp.indices()[0]
                t[idx] = adam1(t[idx], gsp.values())
                t.grad = None
            for t in dense_p:
                if t.grad is not None:
                    t.copy_(adam1(t, t.grad))
                    t.grad = None
        return loss

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    for t in (T, Bt, Lt):
        if t is not None:
            t.requires_grad_(False)
    return {"value": round(B / (ms * 1e-3), 1), "unit": "samples/s", "ms_per_step": round(ms, 4),
            "kind": "torch-eager-cuda", "steps": steps,
            "what": "same step with stock PyTorch ops on the same GPU and tables: F.embedding(sparse=True), cuBLAS, "
                    "autograd, sparse-grad coalesce + indexed first-step Adam (non-deterministic atomics allowed)"}


# --------------------------------------------------------------------------------------- CPU oracle arm
def cpu_baseline(args, w, steps=3, warmup=1):
    """The torch-CPU oracle doing the same DeepFM/DCN/xDeepFM fwd+bwd(+sparse update) on a bounded sample."""
    import oracle

    torch.manual_seed(2019)
    m, n_dense = w["m"], w["n_dense"]
    k = args.k or w["k"]
    rows = min(args.cpu_rows, args.rows or w["rows"])
    B = min(args.cpu_batch, args.batch or w["batch"])
    # every host core: torchrun exports OMP_NUM_THREADS=1, which would silently make this a one-core run
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if torch.get_num_threads() != ncpu:
        torch.set_num_threads(ncpu)
    cores = torch.get_num_threads()
    g = torch.Generator().manual_seed(2020)
    table = torch.randn(m * rows, k, generator=g) * 0.01
    offs = np.arange(m + 1, dtype=np.int64) * rows
    bias_t = torch.zeros(m * rows)
    lin_t = torch.zeros(m * rows)
    d = m * k + n_dense
    deep = list(w["deep"])
    dims = [d] + deep
    Ws = [(torch.randn(dims[i], dims[i + 1], generator=g) * 0.05).requires_grad_() for i in range(len(deep))]
    bs = [torch.zeros(dims[i + 1], requires_grad=True) for i in range(len(deep))]
    w_out = (torch.randn(deep[-1], 1, generator=g) * 0.05).requires_grad_()
    w0 = torch.zeros(1, requires_grad=True)
    extra = []
    if w["model"] == "DCN":
        L = w["cross"]
        cw = (torch.randn(L, d, generator=g) * 0.05).requires_grad_()
        cb = torch.zeros(L, d, requires_grad=True)
        co = (torch.randn(d, 1, generator=g) * 0.05).requires_grad_()
        c0 = torch.zeros(1, requires_grad=True)
        extra = [cw, cb, co, c0]
    if w["model"] == "xDeepFM":
        shapes, final = oracle.cin_layer_shapes(m, w["cin"])
        filt = [(torch.randn(*s, generator=g) * 0.05).requires_grad_() for s in shapes]
        fb = [torch.zeros(s[-1], requires_grad=True) for s in shapes]
        cin_w = (torch.randn(final, 1, generator=g) * 0.05).requires_grad_()
        cin_w0 = torch.zeros(1, requires_grad=True)
        extra = filt + fb + [cin_w, cin_w0]
    dense_params = Ws + bs + [w_out, w0] + extra
    lr = 1e-3

    def step(i):
        ids, dense, y = make_batch(7 + i, B, m, rows, n_dense, args.ids)
        keys = oracle.global_rows(ids, offs)
        kt = torch.from_numpy(keys)
        e = table[kt].requires_grad_()  # [B, m, k] gathered rows (A1-A3)
        bias = bias_t[kt].unsqueeze(-1).requires_grad_()
        linv = lin_t[kt].requires_grad_()
        dn = torch.from_numpy(dense)
        lin_logit = linv.sum(1, keepdim=True)
        if w["model"] == "DeepFM":
            logit = oracle.deepfm_logit(e, bias, lin_logit, dn, (Ws, bs, w_out, w0), oracle.relu)
        elif w["model"] == "DCN":
            logit = oracle.dcn_logit(e, lin_logit, dn, (Ws, bs, w_out, w0), tuple(extra), oracle.relu)
        else:
            nl = len(w["cin"])
            logit = oracle.xdeepfm_logit(e, lin_logit, dn, (Ws, bs, w_out, w0), (extra[:nl], extra[nl:2 * nl], extra[-2], extra[-1]))
        loss = oracle.create_loss(torch.from_numpy(y), oracle.prediction(logit), "classification")
        loss.backward()
        # A4: deterministic segment sum of the embedding gradient, then the fresh-optimizer update on touched rows
        uniq, sums, _, _ = oracle.segment_sum_sorted(keys.reshape(-1), e.grad.reshape(-1, k).numpy())
        ut = torch.from_numpy(uniq)
        table[ut] = oracle.fresh_optimizer_step(table[ut], torch.from_numpy(sums), "adam", lr)
        with torch.no_grad():
            for p in dense_params:
                if p.grad is not None:
                    p.copy_(oracle.fresh_optimizer_step(p, p.grad, "adam", lr))
                    p.grad = None
        return loss.detach().item()

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    dt = (time.perf_counter() - t0) / steps
    return {"value": round(B / dt, 1), "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps of B={B} (same model; tables {m} x {rows} x k={k} to bound host RAM), "
                      f"torch-CPU oracle fwd+bwd + numpy segment-sum + sparse update, {dt * 1e3:.1f} ms/step",
            "cpu_model": _cpu_model(), "ms_per_step": round(dt * 1e3, 2), "batch": B, "rows_per_table": rows}


def _cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    steps = max(1, min(args.steps, 10))
    cpu = cpu_baseline(args, w, steps=steps, warmup=max(1, min(args.warmup, 2)))
    k = args.k or w["k"]
    out = {
        "impl": "reference", "metric": "CTR train samples/sec", "value": cpu["value"], "unit": "samples/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": args.warmup,
        "ms_per_step": cpu["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        # the configuration this arm ACTUALLY runs: a bounded sample of the workload (smaller batch and tables so that
        # the CPU run ends within minutes and fits host RAM); "full_workload" names what it is a sample of
        "config": {"workload": w["desc"] + " - CPU sample", "model": w["model"], "fields": w["m"],
                   "rows_per_table": cpu["rows_per_table"], "k": k, "n_dense": w["n_dense"], "batch_per_gpu": cpu["batch"],
                   "ids": args.ids, "host_threads": cpu["cores"],
                   "full_workload": {"rows_per_table": args.rows or w["rows"], "batch_per_gpu": args.batch or w["batch"]}},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
