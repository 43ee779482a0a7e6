#!/bin/bash
# A/B of two builds of libb200vf.so on ONE box (the pool's boxes differ by +-10 % in sustained clocks): alternates
# ab/<name>.so into place and runs the bench for each.   tools/ab_libs.sh prev new [reps]
set -u
cd "$(dirname "$0")/.."
A=$1; B=$2; REPS=${3:-2}
mkdir -p gpurun_out
cp voicefixer_main_b200/libb200vf.so /tmp/keep.so
for rep in $(seq $REPS); do
  for v in $A $B; do
    cp ab/$v.so voicefixer_main_b200/libb200vf.so
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/op_${v}_$rep.json 2> gpurun_out/bench_${v}_$rep.err | tail -1 > gpurun_out/bench_${v}_$rep.json
    python -c "
import json;d=json.load(open('gpurun_out/bench_${v}_$rep.json'));print('$v', round(d['value'],1), round(d['e2e']['value'],1), {k:round(x,2) for k,x in d['stage_ms'].items()}, d['parity']['wav_rms'], d['clocks']['sm_mhz'])"
  done
done
cp /tmp/keep.so voicefixer_main_b200/libb200vf.so
