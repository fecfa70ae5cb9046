#!/bin/bash
# Per-kernel ncu metrics of the bucket stage (affine rounds + serial tail) of one default prove() at n = 2^16:
#   bash tools/ncu_affine_stage.sh <tag>      -> gpurun_out/<tag>_aff_metrics.csv, gpurun_out/<tag>_affine_stage.md / _affine_traffic.json
tag=${1:-r02x}
ncu --metrics gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size \
    --clock-control none -k regex:"k_aff|k_msm_accumulate|k_msm_fixup|k_msm_heavy" --launch-count 120 --csv --log-file gpurun_out/${tag}_aff_metrics.csv \
    timeout 300 python tools/profile_prove.py 16 2 > gpurun_out/${tag}_aff.log 2>&1
python - "$tag" <<'PY'
import csv, json, re, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/{tag}_aff_metrics.csv")) if len(r) > 10]
I = {h: i for i, h in enumerate(rows[0])}
UN = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0}
agg = {}
for r in rows[1:]:
    k = (int(r[I["ID"]]), re.sub(r"\(.*", "", r[I["Kernel Name"]]).replace("void ", ""))
    agg.setdefault(k, {})[r[I["Metric Name"]]] = float(r[I["Metric Value"]].replace(",", "")) * UN.get(r[I["Metric Unit"]], 1.0)
ids = sorted(agg)
firsts = [i for i in ids if i[1].startswith("k_aff_prepare_first")]
last = [i for i in ids if i[0] >= firsts[-1][0]] if firsts else ids      # the stage of the last proof
lines = [f"# {tag}: the bucket stage of one default prove() at n = 2^16, kernel by kernel (ncu, `--clock-control none`)", "",
         "| # | kernel | ms | fmaheavy % | DRAM read GB | DRAM write GB | warps active % | regs | grid |", "|---:|---|---:|---:|---:|---:|---:|---:|---:|"]
tot_ms = tot_b = 0.0
for n, i in enumerate(last):
    m = agg[i]
    ms, rd, wr = m["gpu__time_duration.sum"], m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
    tot_ms += ms; tot_b += rd + wr
    lines.append(f"| {n} | `{i[1]}` | {ms:.3f} | {m['sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f} | {rd / 1e9:.3f} | {wr / 1e9:.3f} | "
                 f"{m['sm__warps_active.avg.pct_of_peak_sustained_active']:.1f} | {int(m['launch__registers_per_thread'])} | {int(m['launch__grid_size'])} |")
lines += ["", f"Stage total: {tot_ms:.2f} ms summed under the profiler, {tot_b / 1e9:.2f} GB of DRAM traffic (read + write)."]
open(f"gpurun_out/{tag}_affine_stage.md", "w").write("\n".join(lines) + "\n")
json.dump({"kernel": "affine bucket stage (all kernels of one proof)", "dram_bytes_per_launch": tot_b, "ms_under_profiler": tot_ms, "source": f"{tag}_aff_metrics.csv"},
          open(f"gpurun_out/{tag}_affine_traffic.json", "w"))
print("\n".join(lines[-14:]))
PY
