#!/bin/bash
# `ncu --set full` of the first- and second-round kernels of the affine bucket stage (one default prove() at n = 2^16):
#   bash tools/ncu_affine_full.sh <tag>   -> gpurun_out/<tag>_aff_full_raw.csv (raw page; the .ncu-rep stays in /tmp on the box)
tag=${1:-r02x}
ncu --set full --clock-control none --import-source on -k regex:"k_aff_prefix|k_aff_add" --launch-count 4 -o /tmp/${tag}_aff_full -f \
    timeout 300 python tools/profile_prove.py 16 1 > gpurun_out/${tag}_aff_full.log 2>&1
ncu -i /tmp/${tag}_aff_full.ncu-rep --page raw --csv > gpurun_out/${tag}_aff_full_raw.csv 2>> gpurun_out/${tag}_aff_full.log
ls -la /tmp/${tag}_aff_full.ncu-rep gpurun_out/${tag}_aff_full_raw.csv
