#!/usr/bin/env python
"""Digest gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and an `ncu --set full` report into small,
committed summaries under profiles/.   usage: ncu_summary.py <tag> [launches.csv] [report.ncu-rep]"""
import collections
import csv
import io
import os
import json
import subprocess
import sys

tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv"
rep = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/prof_top.ncu-rep"

rows = list(csv.DictReader(l for l in open(launches) if l.startswith('"')))
n_all = len(rows)
if os.environ.get("SKIP_SETUP", "1") == "1":
    # model construction (random init of the tables: torch elementwise kernels over 66 GB) precedes the first step;
    # the list starts at the first own kernel
    first = next((i for i, r in enumerate(rows) if "rm::" in r["Kernel Name"][:40]), 0)
    rows = rows[first:]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r["Kernel Name"], [0, 0.0, r["Grid Size"], r["Block Size"]])
    a[0] += 1
    a[1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_launches.md", "w") as f:
    f.write(f"# ncu launch list `{tag}` — {len(rows)} launches from the first own kernel on ({n_all} captured incl. model "
            f"construction), {tot:.1f} us of kernel time (serialised, cold cache: shares, not absolutes)\n\n")
    f.write("| us total | launches | share | grid | block | kernel |\n|---:|---:|---:|---|---|---|\n")
    for n, (c, t, g, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
        own = "**" if n.startswith("rm::") or "rm::" in n[:30] else ""
        f.write(f"| {t:.1f} | {c} | {100 * t / tot:.1f}% | {g} | {b} | {own}`{n[:110]}`{own} |\n")
    own_t = sum(t for n, (c, t, g, b) in agg.items() if "rm::" in n[:30])
    f.write(f"\nown kernels (`rm::`): {own_t:.1f} us = {100 * own_t / tot:.1f}% of kernel time\n")
print(f"profiles/{tag}_launches.md")

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "sm__cycles_elapsed.max",
]
try:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    h, units = r[0], r[1]
    res = []
    for row in r[2:]:
        d = {"kernel": row[h.index("Kernel Name")][:80]}
        for w in WANT:
            if w in h:
                d[w] = f"{row[h.index(w)]} {units[h.index(w)]}".strip()
        # tensor-pipe metrics: keep anything mentioning tensor
        for i, name in enumerate(h):
            if "tensor" in name and ".avg.pct" in name and name not in d and row[i] not in ("0", "", "n/a"):
                d[name] = f"{row[i]} {units[i]}".strip()
        res.append(d)
    json.dump(res, open(f"profiles/{tag}_ncu_full.json", "w"), indent=1)
    print(f"profiles/{tag}_ncu_full.json", len(res))
except Exception as e:  # noqa
    print("no full report:", e)
