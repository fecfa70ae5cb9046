#!/usr/bin/env python3
"""Turns the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_profiles.py <tag> <launches.csv> <kernel.ncu-rep | -> [first_launch_id_of_last_proof]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
first = int(sys.argv[4]) if len(sys.argv) > 4 else None
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]
ki, vi, ui, idi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
data = [(int(r[idi]), re.sub(r"\(.*", "", r[ki]).replace("void ", ""), float(r[vi].replace(",", ""))) for r in rows[1:]]
unit = rows[1][ui]
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1e-6)


def table(ds):
    agg = collections.OrderedDict()
    for _, k, v in ds:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    lines = ["| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {c} | {v:.3f} | {100 * v / tot:.1f}% |")
    return tot, "\n".join(lines)


ids = [d[0] for d in data]
# the last proof = launches from the last k_prove_points-1 (fr_to_mont precedes it) to the end
if first is None:
    pp = [d[0] for d in data if d[1].startswith("k_prove_points")]
    first = pp[-1] - 2 if pp else ids[0]
proof = [d for d in data if d[0] >= first]
setup = [d for d in data if d[0] < (min(d[0] for d in data if d[1].startswith("k_prove_points")) - 2 if any(d[1].startswith("k_prove_points") for d in data) else first)]
tot_p, tab_p = table(proof)
tot_s, tab_s = table(setup)
md = [f"# {tag}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)", "",
      "Per-launch times are cold-cache and serialised by the profiler: compare SHARES, not absolutes.",
      f"Source: `{os.path.basename(launches)}` ({len(data)} launches).", "",
      f"## One prove() at n=2^16, Q=8 (last proof of the run): {len(proof)} launches, {tot_p:.2f} ms summed", "", tab_p, "",
      f"## SRS.new + circuit load: {len(setup)} launches, {tot_s:.2f} ms summed", "", tab_s, ""]
open(os.path.join(out_dir, f"{tag}_launches.md"), "w").write("\n".join(md))

if rep == "-":   # launch list only
    print("wrote", f"{tag}_launches.md", "proof launches", len(proof), "summed ms", round(tot_p, 2))
    sys.exit(0)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]
m = {k: (val, un) for k, un, val in zip(h, u, v)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
lines = [f"# {tag}: `ncu --set full --clock-control none --import-source on` of the dominant kernel", "",
         f"Source: `{os.path.basename(rep)}` (kept in gpurun_out/, not tracked: 22 MB).", "", "| metric | value | unit |", "|---|---:|---|"]
for k in keys:
    if k in m:
        lines.append(f"| `{k}` | {m[k][0]} | {m[k][1]} |")


def num(k):
    val, un = m[k]
    x = float(val.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(un, 1.0)


traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
lines += ["", f"DRAM traffic per launch = {traffic / 1e9:.3f} GB (read + write)."]
open(os.path.join(out_dir, f"{tag}_accumulate_ncu.md"), "w").write("\n".join(lines) + "\n")
json.dump({"kernel": m["Kernel Name"][0], "dram_bytes_per_launch": traffic, "source": os.path.basename(rep)},
          open(os.path.join(out_dir, f"{tag}_accumulate_traffic.json"), "w"))
print("wrote profiles for", tag, "proof launches", len(proof), "traffic GB", traffic / 1e9)
