timeout 700 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -4
bash tools/run_profile.sh 2>&1 | tail -12
