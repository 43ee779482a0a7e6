"""Run a few restore() steps of the benchmark workload - the command ncu wraps (see profiles/README.md)."""
import argparse
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
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
model = VoiceFixer().load_state_dict(make_state(1234)).eval().to("cuda:0")
wav = synth_batch(a.batch, int(a.seconds * 44100), 1000).cuda()
out = torch.empty_like(wav)
for _ in range(a.steps):
    model.restore(wav, out)
torch.cuda.synchronize()
model._engine().check_errors()
print("launches", model._engine().launch_count())
