for spec in enc3_b2_conv1:17 enc1_b2_conv1:1 voc_res1_1_a:131; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s $((179+idx)) -c 1 -f -o gpurun_out/prof_$name python tools/profile_step.py --steps 2 > gpurun_out/ncu_$name.log 2>&1; tail -1 gpurun_out/ncu_$name.log | cut -c1-200
done
