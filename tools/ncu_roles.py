"""Warp-role view of an ncu report of gemm_tc_kernel: samples in the producer / epilogue / MMA code regions and the
share each role spends polling its barriers.   python tools/ncu_roles.py report.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, body = rows[1], rows[2:]
isrc, ism, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
tma = [i for i, r in enumerate(body) if "UTMALDG" in r[isrc]]
mma = [i for i, r in enumerate(body) if "UTCHMMA" in r[isrc]]
ldtm = [i for i, r in enumerate(body) if "LDTM" in r[isrc]]
# region borders: producer = [0, first line after the last TMA whose exec count drops), MMA = from the wait before the first UTCHMMA
exec_tma = int(body[tma[-1]][ie])
p_end = max(tma) + 1
while p_end < len(body) and p_end < min(ldtm) and int(body[p_end][ie]) * 8 >= exec_tma * 0 + 1 and "LDTM" not in body[p_end][isrc] and p_end < max(tma) + 60:
    p_end += 1
m_start = min(mma)
while m_start > max(ldtm) and "TRYWAIT" not in body[m_start][isrc]:
    m_start -= 1
m_start -= 60
tot = sum(int(r[ism]) for r in body)


def region(a, b, name):
    s = sum(int(body[i][ism]) for i in range(a, b))
    polls = 0
    for i in range(a, b):
        if "TRYWAIT" in body[i][isrc]:
            polls += int(body[i][ism]) + int(body[i + 1][ism])
    print(f"{name:10s} lines [{a},{b})  samples {s:6d} ({100.0 * s / tot:5.1f}% of all)   polling {polls:6d} ({100.0 * polls / max(s, 1):5.1f}% of the role)")


print(rows[0][1] if len(rows[0]) > 1 else "", "total samples", tot)
region(0, p_end, "producer")
region(p_end, m_start, "epilogue")
region(m_start, len(body), "mma+exit")
