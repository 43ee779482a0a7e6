timeout 700 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -5
bash tools/run_ab.sh nbuf4 "VF_X=0" nbuf2 "VF_TUNE_NBUF=2" nbuf4b "VF_X=0" nbuf2b "VF_TUNE_NBUF=2"
