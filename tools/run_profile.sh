#!/bin/bash
# Runs on the GPU box (gpurun): bench line + ncu launch list of the same command + ncu --set full captures.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python bench.py --steps 3 --warmup 3 --profile-out gpurun_out/op_profile.json 2>&1 | tail -1 > gpurun_out/bench_r1.json
cut -c1-300 gpurun_out/bench_r1.json
# launch list of one timed step (3 warm-up steps x 193 launches skipped)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 579 -c 193 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
# full captures: index = position among the 179 tcgen05 GEMM launches of one restore (second restore profiled)
for spec in enc1_b2_conv1:1 enc3_b2_conv1:17 voc_res0_1_a:114 voc_res1_1_b:132 voc_res3_1_a:165 voc_res3_1_b:166; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s $((179+idx)) -c 1 -f \
      -o gpurun_out/prof_$name python tools/profile_step.py --steps 2 > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-120
done
ls gpurun_out/
