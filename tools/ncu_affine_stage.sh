ncu --metrics gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size --clock-control none -k regex:"k_aff|k_msm_accumulate|k_msm_fixup|k_msm_heavy" -c 40 --csv --log-file gpurun_out/r02p_aff_metrics.csv timeout 300 python tools/time_prove_modes.py acc_mode=3 > gpurun_out/r02p.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02p_aff_metrics.csv")) if len(r)>10]
hdr=rows[0]; I={h:i for i,h in enumerate(hdr)}
agg={}
for r in rows[1:]:
    key=(r[I["ID"]], r[I["Kernel Name"]][:60])
    agg.setdefault(key,{})[r[I["Metric Name"]]]=r[I["Metric Value"]]+" "+r[I["Metric Unit"]]
for k,v in agg.items():
    print(k[0],k[1],"|",v.get("gpu__time_duration.sum"),"| fma",v.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),"| rd",v.get("dram__bytes_read.sum"),"| wr",v.get("dram__bytes_write.sum"),"| warps",v.get("sm__warps_active.avg.pct_of_peak_sustained_active"),"| regs",v.get("launch__registers_per_thread"),"| grid",v.get("launch__grid_size"))
PY
