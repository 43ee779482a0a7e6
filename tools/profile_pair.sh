#!/bin/bash
# ncu --set full capture of one fused residual-pair launch (voc.res3.1.pair) + the raw and source pages as CSV.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_tc_kernel -s 9 -c 1 -f \
    -o gpurun_out/prof_voc_res3_1_pair python tools/profile_step.py --steps 2 > gpurun_out/ncu_pair.log 2>&1
tail -2 gpurun_out/ncu_pair.log | cut -c1-160
ncu -i gpurun_out/prof_voc_res3_1_pair.ncu-rep --page raw --csv > gpurun_out/prof_voc_res3_1_pair_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_voc_res3_1_pair.ncu-rep --page source --csv 2>/dev/null | cut -c1-600 > gpurun_out/prof_voc_res3_1_pair_source.csv
ls -la gpurun_out/prof_voc_res3_1_pair*
