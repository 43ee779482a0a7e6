timeout 700 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -5
bash tools/run_ab.sh base "VF_X=0" c32_2 "VF_TUNE_CTAS_32_3=2" base2 "VF_X=0"
for spec in enc1_b2_conv1:1 voc_res1_1_b:132 voc_res3_1_a:165; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s $((179+idx)) -c 1 -f -o gpurun_out/prof_$name python tools/profile_step.py --steps 2 > gpurun_out/ncu_$name.log 2>&1; tail -1 gpurun_out/ncu_$name.log | cut -c1-200
done
