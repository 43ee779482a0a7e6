"""Turn the ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01

Inputs (produced on the GPU box by tools/run_profile.sh):
  gpurun_out/launches.csv           ncu --metrics gpu__time_duration.sum launch list of one bench.py step
  gpurun_out/prof_<name>.ncu-rep    ncu --set full captures of selected GEMM launches
  gpurun_out/op_profile.json        CUDA-event per-launch profile written by bench.py --profile-out
  gpurun_out/bench_r1.json          the bench line of the same build
"""
import collections
import csv
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = [f"# ncu / event profile summary {tag}", ""]

# ---- launch list -------------------------------------------------------------------------------------------
lp = os.path.join(G, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if r == hdr or len(r) <= iv:
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        us = v / 1e3 if r[iu] == "ns" else (v * 1e3 if r[iu] == "ms" else v)   # ncu prints ns / us / ms
        name = re.sub(r"\(.*", "", r[ik]).replace("void vf::", "").replace("vf::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    out += ["## Launch list of one bench.py step (ncu --metrics gpu__time_duration.sum --clock-control none; serialised, cold cache: compare shares)",
            "", f"{sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms summed", "", "| kernel | launches | ms | share |", "|---|---|---|---|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {us / 1e3:.3f} | {100 * us / tot:.1f}% |")
    out.append("")
    with open(os.path.join(P, f"{tag}_launches.csv"), "w") as f:
        f.write(open(lp, errors="ignore").read())

# ---- full captures -----------------------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]
reps = sorted(set(glob.glob(os.path.join(G, "prof_*.ncu-rep")) + glob.glob(os.path.join(G, "prof_*_raw.csv"))))
seen = set()
if reps:
    out += ["## ncu --set full captures (one launch each; `traffic` = dram read + write)", ""]
    table = []
    for rep in reps:
        name = os.path.basename(rep)[5:-8]
        if name in seen:
            continue
        seen.add(name)
        if rep.endswith(".csv"):      # exported on the GPU box by tools/run_profile.sh (the binary report stays there)
            txt = open(rep, errors="ignore").read()
        else:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(txt.splitlines()))
        if len(rr) < 3:
            continue
        hdr, unit, val = rr[0], rr[1], rr[2]
        d = {h: (val[i], unit[i]) for i, h in enumerate(hdr)}
        table.append((name, d))
        with open(os.path.join(P, f"{tag}_{name}_raw.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit", "value"])
            for h in hdr:
                if any(k in h for k in ("dram__", "lts__t", "lts__throughput", "tensor", "sm__throughput", "warps_active", "launch__", "gpu__time",
                                        "xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum", "smsp__cycles_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")):
                    w.writerow([h, d[h][1], d[h][0]])
    # label -> dram traffic of the captured launch, for bench.py's roofline.traffic
    def _bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    traffic = {}
    for name, d in table:
        label = name.replace("voc_res", "voc.res").replace("_conv", ".conv")
        parts = name.split("_")
        if name.startswith("voc_res"):
            label = f"voc.res{parts[1][3:]}.{parts[2]}.{parts[3]}"
        else:
            label = f"{parts[0]}.{parts[1]}.{parts[2]}"
        traffic[label] = {"kernel": d["Kernel Name"][0], "dram_bytes": _bytes(*d["dram__bytes_read.sum"]) + _bytes(*d["dram__bytes_write.sum"]),
                          "time_us": d["gpu__time_duration.sum"][0] + " " + d["gpu__time_duration.sum"][1]}
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    out += ["| metric | " + " | ".join(n for n, _ in table) + " |", "|---|" + "---|" * len(table)]
    out.append("| kernel | " + " | ".join(re.sub(r"\(.*", "", d["Kernel Name"][0]).replace("void ", "") for _, d in table) + " |")
    for m in WANT:
        out.append(f"| {m} | " + " | ".join(f"{d.get(m, ('', ''))[0]} {d.get(m, ('', ''))[1]}" for _, d in table) + " |")
    out.append("")

# ---- event profile -----------------------------------------------------------------------------------------
op = os.path.join(G, "op_profile.json")
if os.path.exists(op):
    d = json.load(open(op))
    ops = d["ops"]
    tot = sum(o["ms"] for o in ops)
    g = collections.OrderedDict()
    for o in ops:
        l = o["label"]
        key = l.split(".")[0]
        if l.startswith("voc.res"):
            key = ".".join(l.split(".")[:2]) + "." + l.split(".")[-1]
        elif l.startswith("voc."):
            key = l
        e = g.setdefault(key, [0.0, 0.0, 0.0, 0, set()])
        e[0] += o["ms"]; e[1] += o["flops"]; e[2] += o["bytes"]; e[3] += 1; e[4].add((o["bn"], o["bk"]))
    out += ["## CUDA-event per-launch profile of one restore() step (bench.py --profile-out; B = 32 x 10 s)", "",
            f"stage ms: {d['stage_ms']}; summed launches {tot:.2f} ms", "",
            "| group | launches | ms | algorithmic TFLOP/s | min-traffic GB/s | tile (BN,BK) |", "|---|---|---|---|---|---|"]
    for k, (ms, fl, by, n, cfg) in g.items():
        if ms > 0.05:
            out.append(f"| {k} | {n} | {ms:.3f} | {fl / ms / 1e9:.1f} | {by / ms / 1e6:.0f} | {sorted(cfg)} |")
    out.append("")
    json.dump(d, open(os.path.join(P, f"{tag}_op_profile.json"), "w"))
bp = os.path.join(G, "bench_r1.json")
if os.path.exists(bp):
    out += ["## bench.py line of the same build", "", "```json", open(bp).read().strip(), "```", ""]
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(out))
print("\n".join(out[:60]))
