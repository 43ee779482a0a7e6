#!/bin/bash
# A/B of tuning knobs on the GPU box: parity subset + one bench (no CPU baseline) + per-launch profile per variant.
#   bash tools/run_ab.sh name1 "ENV1=.. ENV2=.." name2 "..." ...
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  name=$1; envs=$2; shift 2
  par=$(env $envs timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=line -k "golden or vocoder_vs or tcgen05" 2>&1 | tail -1)
  env $envs timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/ab_$name.json 2>&1 | tail -1 > gpurun_out/ab_$name.line
  python - "$name" "$par" <<'PY'
import json, sys
name, par = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/ab_{name}.line").read())
    print(name, "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "stage_ms", {k: round(v, 2) for k, v in d["stage_ms"].items()}, "| parity:", par)
except Exception as ex:
    print(name, "FAILED", ex, open(f"gpurun_out/ab_{name}.line").read()[:300], "| parity:", par)
PY
done
