"""A/B of the host-buffer entry point: pipelined (internal copy / compute streams) vs single stream, alternating in ONE process so
clock drift hits both arms alike.   python tools/ab_host_pipeline.py [--steps 10] [--reps 4]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_batch  # noqa: E402
from voicefixer_main_b200 import VoiceFixer  # noqa: E402
from voicefixer_main_b200.weights import make_state  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
model = VoiceFixer().load_state_dict(make_state(1234)).eval().to("cuda:0")
eng = model._engine()
n = int(a.seconds * 44100)
host_in = synth_batch(a.batch, n, 1000).pin_memory()
host_out = torch.empty_like(host_in).pin_memory()
dev_in, dev_out = host_in.cuda(), torch.empty(a.batch, n, device="cuda")


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps


for _ in range(3):
    model.restore(dev_in, dev_out)
    model.restore_host(host_in, host_out)
res = {"device_resident": [], "host_pipelined": [], "host_single_stream": []}
for r in range(a.reps):
    res["device_resident"].append(timed(lambda: model.restore(dev_in, dev_out)))
    eng.set_option("host_pipeline", 1)
    model.restore_host(host_in, host_out)
    res["host_pipelined"].append(timed(lambda: model.restore_host(host_in, host_out)))
    ref = host_out.clone()
    eng.set_option("host_pipeline", 0)
    model.restore_host(host_in, host_out)
    res["host_single_stream"].append(timed(lambda: model.restore_host(host_in, host_out)))
    assert torch.equal(ref, host_out)
eng.set_option("host_pipeline", 1)
eng.check_errors()
print(json.dumps({k: {"ms_per_step": v, "mean": sum(v) / len(v)} for k, v in res.items()}))
