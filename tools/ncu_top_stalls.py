"""Top stall locations (SASS) of an ncu report: python tools/ncu_top_stalls.py report.ncu-rep [N]"""
import csv
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
i_src, i_smp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = rows[2:]
tot = sum(int(r[i_smp]) for r in body)
print(rows[0][1], "total samples", tot)
order = sorted(range(len(body)), key=lambda i: -int(body[i][i_smp]))[:n]
for i in sorted(order):
    r = body[i]
    ctx = " | ".join(body[j][i_src].strip()[:28] for j in range(max(0, i - 2), i))
    print(f"{100.0 * int(r[i_smp]) / tot:5.1f}%  exec={r[i_exec]:>9s}  {r[i_src].strip()[:70]:70s}  <- {ctx}")
