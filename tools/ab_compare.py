"""Compare per-launch profiles written by tools/run_ab.sh: python tools/ab_compare.py base other [other...]"""
import json
import re
import sys
from collections import OrderedDict


def groups(path):
    prof = json.load(open(path))
    ops = prof["ops"] if isinstance(prof, dict) else prof
    g = OrderedDict()
    for o in ops:
        lab = o.get("label", "?")
        key = re.sub(r"\.\d+\.([ab])$", r".\1", lab)          # voc.res3.5.a -> voc.res3.a
        key = key.split(".")[0] if key.startswith(("enc", "dec", "post", "bottleneck")) else key
        g[key] = g.get(key, 0.0) + o["ms"]
    return g


names = sys.argv[1:]
gs = [groups(f"gpurun_out/ab_{n}.json") for n in names]
print(f"{'group':16s}" + "".join(f"{n:>12s}" for n in names))
for k in gs[0]:
    print(f"{k:16s}" + "".join(f"{g.get(k, float('nan')):12.3f}" for g in gs))
print(f"{'TOTAL':16s}" + "".join(f"{sum(g.values()):12.3f}" for g in gs))
