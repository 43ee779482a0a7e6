#!/bin/bash
# A/B of environment knobs on ONE box:  tools/ab_env.sh "A=1" "B=2 C=3" ...   ("-" = no override); two alternating repetitions
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
  i=0
  for spec in "$@"; do
    i=$((i+1))
    envs=""; [ "$spec" != "-" ] && envs="$spec"
    env $envs python bench.py --steps 8 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/op_v${i}_$rep.json 2> gpurun_out/bench_v${i}_$rep.err | tail -1 > gpurun_out/bench_v${i}_$rep.json
    python -c "
import json;d=json.load(open('gpurun_out/bench_v${i}_$rep.json'));print('[$spec]', round(d['value'],1), round(d['e2e']['value'],1), {k:round(x,2) for k,x in d['stage_ms'].items()}, d['parity']['wav_rms'], d['clocks']['sm_mhz'])"
  done
done
