"""Summarise an ncu report (gpurun_out/*.ncu-rep) and a launch list into profiles/.
usage: python tools/profile_summary.py <rep> <launches.csv> <tag>"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out_md = os.path.join(ROOT, "profiles", f"{tag}_summary.md")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
]
lines = [f"# ncu --set full summary ({tag})", "", f"source: `{os.path.basename(rep)}` (ncu --set full --clock-control none --import-source on)", ""]
traffic = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    base = name.split("<")[0]
    lines.append(f"## {name}")
    lines.append("")
    lines.append("| metric | value | unit |")
    lines.append("|---|---|---|")
    for k in KEYS:
        if k in idx:
            lines.append(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
    rd = float(r[idx["dram__bytes_read.sum"]]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[idx["dram__bytes_write.sum"]]]
    dur = float(r[idx["gpu__time_duration.sum"]])
    traffic.setdefault(base, rd + wr)
    lines.append(f"| DRAM traffic per launch | {rd + wr:.4g} | byte |")
    stalls = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                v = float(r[idx[h]])
            except ValueError:
                continue
            if v > 0.25:
                stalls.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    lines.append("")
    lines.append("top stall reasons (warps per issue-active): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:6]))
    lines.append("")

# launch list: share of the step per kernel
agg = defaultdict(lambda: [0, 0.0])
with open(launches) as f:
    rd = csv.reader(l for l in f if not l.startswith("=="))
    h = next(rd)
    ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    ui = h.index("Metric Unit")
    for r in rd:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3}.get(r[ui], 1e-6)
        n = r[ki].split("(")[0].replace("void ", "")
        agg[n][0] += 1
        agg[n][1] += v * scale
tot = sum(v[1] for v in agg.values())
lines += ["## launch list (ncu --metrics gpu__time_duration.sum, cold-cache, serialised: compare SHARES)", "",
          f"source: `{os.path.basename(launches)}`; total {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches", "",
          "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    lines.append(f"| {n[:90]} | {c} | {ms:.3f} | {ms / tot * 100:.1f}% |")
open(out_md, "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(out_md)
