for seg in 16 24 32 48; do
  echo "== SEG_MMAS=$seg"
  VF_TUNE_SEG_MMAS=$seg timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s --tb=line -k "unet_matches_reference_golden" 2>&1 | grep -E "log-mel|passed|failed"
done
bash tools/run_ab.sh seg16 "VF_TUNE_SEG_MMAS=16" seg24 "VF_TUNE_SEG_MMAS=24" seg32 "VF_TUNE_SEG_MMAS=32" seg48 "VF_TUNE_SEG_MMAS=48" seg16b "VF_TUNE_SEG_MMAS=16"
