timeout 700 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -8
bash tools/run_ab.sh base "VF_X=0" c32_2 "VF_TUNE_CTAS_32_3=2" c64_1 "VF_TUNE_CTAS_64_3=1" halo2 "VF_TUNE_HALO=2" halo2c "VF_TUNE_HALO=2 VF_TUNE_CTAS_32_3=2"
