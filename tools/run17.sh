timeout 900 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -6
bash tools/run_ab.sh base "VF_X=0" base2 "VF_X=0"
