#!/usr/bin/env python3
"""profiles/<tag>_fr_ncu.md from an `ncu --set full` report of the Fr-side and SRS kernels (tools/profile_fr.py):
HBM GB/s of every launch against the measured copy bandwidth, multiplier-pipe and SM utilisation.
  python tools/summarize_fr.py <tag> <report.ncu-rep>"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}
try:
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    src = "MEASURED_PEAKS.json (of measured)"
except Exception:
    hbm, src = 6650.0, "B200_PROFILING.md fallback (of fallback)"


def val(r, k):
    x = float(r[col[k]].replace(",", ""))
    u = units[col[k]]
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)


agg = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")
    a = agg.setdefault(name, dict(n=0, t=0.0, rd=0.0, wr=0.0, sm=[], fma=[], warps=[], regs=r[col["launch__registers_per_thread"]],
                                  grid=[], l2=[]))
    a["n"] += 1
    a["t"] += val(r, "gpu__time_duration.sum")
    a["rd"] += val(r, "dram__bytes_read.sum")
    a["wr"] += val(r, "dram__bytes_write.sum")
    a["sm"].append(float(r[col["sm__throughput.avg.pct_of_peak_sustained_elapsed"]]))
    a["fma"].append(float(r[col["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]]))
    a["warps"].append(float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]))
    a["l2"].append(float(r[col["lts__t_sector_hit_rate.pct"]]))
    a["grid"].append(r[col["launch__grid_size"]])
avg = lambda v: sum(v) / len(v)
lines = [f"# {tag}: `ncu --set full --clock-control none` of the Fr-side and SRS kernels (one prove() at n = 2^16, SRS.new without tables)", "",
         f"HBM peak = {hbm:.0f} GB/s, {src}.  DRAM GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration, summed over the launches of a kernel.",
         "Per-launch times under ncu are cold-cache and serialised.", "",
         "| kernel | launches | ms | grid | regs | DRAM read | DRAM write | DRAM GB/s | of HBM peak | L2 hit % | SM throughput % | fmaheavy % | warps active % |",
         "|---|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for k, a in agg.items():
    gbs = (a["rd"] + a["wr"]) / a["t"] / 1e9
    lines.append(f"| `{k}` | {a['n']} | {a['t'] * 1e3:.3f} | {'/'.join(sorted(set(a['grid']), key=int))} | {a['regs']} | {a['rd'] / 1e6:.1f} MB | {a['wr'] / 1e6:.1f} MB | "
                 f"{gbs:.0f} | {gbs / hbm:.3f} | {avg(a['l2']):.0f} | {avg(a['sm']):.1f} | {avg(a['fma']):.1f} | {avg(a['warps']):.1f} |")
lines += ["", "Reading: none of these kernels is HBM-bound at these sizes.  `k_fixed_base` and `k_batch_affine` are multiplier-bound (curve arithmetic); `k_pow_tables`, "
          "`k_open_partial`, `k_open_quotient`, `k_build_sxy` spend one or two Fr multiplications per 32-byte element they move (136 LMAC per 32-64 B: the "
          "multiplier pipe, not the 6.5 TB/s, sets their pace); the NTT passes work on a 16.8 MB vector that stays in the 126 MB L2 (DRAM traffic is the "
          "first touch only) and are bound by shared-memory butterflies and launch latency (30-80 us each).", "",
          f"Source: `{os.path.basename(rep)}` (kept in gpurun_out/, not tracked)."]
open(os.path.join(ROOT, "profiles", f"{tag}_fr_ncu.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
