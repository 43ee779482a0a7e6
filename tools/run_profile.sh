#!/bin/bash
# Runs on the GPU box (gpurun): bench line + ncu launch list of the same command + ncu --set full captures.
#   tools/run_profile.sh            -> artefacts under gpurun_out/; then here: python tools/summarize_profiles.py r02
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/op_profile.json 2> gpurun_out/bench_r1.err | tail -1 > gpurun_out/bench_r1.json
cut -c1-300 gpurun_out/bench_r1.json
N=$(python -c "import json;print(json.load(open('gpurun_out/bench_r1.json'))['gpu_launches']//5)")
echo "launches per step: $N"
# launch list of one timed step (3 warm-up steps skipped; the graph's kernel nodes are listed one by one)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*N)) -c $N --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
# full captures: index = position among the tcgen05 GEMM launches of one restore (second restore profiled)
G=$(python - <<'PY'
import json
ops=json.load(open('gpurun_out/op_profile.json'))['ops']
g=[o['label'] for o in ops if o['bn']]
want=['enc1.b2.conv1','enc3.b2.conv1','voc.res0.1.a','voc.res1.1.b','voc.res2.1.a','voc.res2.1.b']
print(len(g), ' '.join(f"{w.replace('.','_')}:{g.index(w)}" for w in want if w in g))
PY
)
NG=${G%% *}; SPECS=${G#* }
echo "gemm launches per step: $NG; captures: $SPECS"
for spec in $SPECS; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s $((NG+idx)) -c 1 -f \
      -o gpurun_out/prof_$name python tools/profile_step.py --steps 2 > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-120
  # gpurun copies back at most 64 MiB: keep the raw / source pages as CSV, and the binary report of two kernels only
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv 2>/dev/null | cut -c1-400 > gpurun_out/prof_${name}_source.csv
  case $name in enc1_b2_conv1) ;; *) rm -f gpurun_out/prof_$name.ncu-rep ;; esac
done
bash tools/profile_pair.sh
du -sh gpurun_out
ls gpurun_out/
