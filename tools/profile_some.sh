#!/bin/bash
# ncu --set full of a few named GEMM launches + the fused pair:  tools/profile_some.sh voc.res2.1.a voc.res2.1.b
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/op_profile.json 2> gpurun_out/bench_r1.err | tail -1 > gpurun_out/bench_r1.json
G=$(WANT="$*" python - <<'PY'
import json, os
ops=json.load(open('gpurun_out/op_profile.json'))['ops']
g=[o['label'] for o in ops if o['bn']]
want=os.environ['WANT'].split()
print(len(g), ' '.join(f"{w.replace('.','_')}:{g.index(w)}" for w in want if w in g))
PY
)
NG=${G%% *}; SPECS=${G#* }
echo "gemm launches per step: $NG; captures: $SPECS"
for spec in $SPECS; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s $((NG+idx)) -c 1 -f \
      -o gpurun_out/prof_$name python tools/profile_step.py --steps 2 > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv 2>/dev/null | cut -c1-400 > gpurun_out/prof_${name}_source.csv
  rm -f gpurun_out/prof_$name.ncu-rep
done
bash tools/profile_pair.sh
rm -f gpurun_out/prof_voc_res3_1_pair.ncu-rep
