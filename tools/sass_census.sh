#!/bin/bash
# Opcode census of the built library: the Blackwell-native tells (tcgen05 MMA / TMEM / TMA / mbarrier) per kernel.
#   tools/sass_census.sh > profiles/r02_sass_census.txt        (runs here: cuobjdump needs no GPU)
set -u
cd "$(dirname "$0")/.."
SO=voicefixer_main_b200/libb200vf.so
echo "# SASS census of $SO ($(stat -c %s $SO) bytes), $(nvcc --version | tail -1)"
echo "# cuobjdump -sass $SO | grep -c <opcode>"
cuobjdump -sass $SO > /tmp/vf_sass.txt
for op in UTCHMMA UTCQMMA UTMALDG UTMASTG LDTM STTM UTCBAR UTCATOMSW SYNCS.ARRIVE SYNCS.PHASECHK ELECT " HMMA" " IMMA" "BRA.U.ANY" " LDL" " STL"; do
  printf "%-16s %d\n" "$op" "$(grep -c -- "$op" /tmp/vf_sass.txt)"
done
echo
echo "# per kernel: UTCHMMA / UTMALDG / LDTM / registers (cuobjdump -res-usage)"
python - <<'PY'
import re, subprocess
txt = open('/tmp/vf_sass.txt').read()
res = subprocess.run(['cuobjdump', '-res-usage', 'voicefixer_main_b200/libb200vf.so'], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r'Function (\S+):', line)
    if m:
        cur = m.group(1)
    m = re.search(r'REG:(\d+)', line)
    if m and cur:
        regs[cur] = int(m.group(1))
dem = {}
names = sorted(set(re.findall(r'Function : (\S+)', txt)))
if names:
    out = subprocess.run(['cu++filt'] + names, capture_output=True, text=True).stdout.splitlines()
    dem = dict(zip(names, out))
parts = re.split(r'\n\s*Function : ', txt)
rows = []
for p in parts[1:]:
    name = p.split('\n', 1)[0].strip()
    rows.append((re.sub(r'\(int\)|\(bool\)', '', dem.get(name, name)).replace('void vf::', '').split('(')[0], p.count('UTCHMMA'), p.count('UTMALDG'), p.count('LDTM'), regs.get(name, -1)))
for r in sorted(rows, key=lambda r: (-r[1], r[0])):
    print(f"{r[0][:70]:70s} UTCHMMA {r[1]:4d}  UTMALDG {r[2]:3d}  LDTM {r[3]:3d}  regs {r[4]}")
PY
