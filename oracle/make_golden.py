"""Generate tests/golden/*.npz from the REFERENCE ITSELF.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):  python oracle/make_golden.py

Every output below is produced by the reference's own modules imported unmodified
(oracle/ref_import.py): FDomainHelper + MelScale (stage A), VoiceFixer.forward ->
Generator -> UNetResComplex_100Mb (stage B), and the handler body of
eval_gsr_voicefixer.py:41-75 (end to end; its vocoder is the restated stage C).
Weights are NOT stored (65 M parameters): they are regenerated from
voicefixer_main_b200.weights.make_state(seed), and a fingerprint of that state is
stored so generator drift is detected instead of silently failing parity.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import, vf_oracle as O            # noqa: E402
from voicefixer_main_b200.weights import make_state      # noqa: E402

SEED = 1234
GOLD = os.path.join(ROOT, "tests", "golden")


def state_fingerprint(sd) -> np.ndarray:
    keys = sorted(k for k in sd if sd[k].is_floating_point())
    probe = [keys[0], keys[len(keys) // 3], keys[len(keys) // 2], keys[-1]]
    return np.array([float(sd[k].double().abs().sum()) for k in probe] +
                    [float(sum(sd[k].double().sum() for k in keys))], dtype=np.float64)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLD, exist_ok=True)
    sd = make_state(SEED)
    model, _ = ref_import.build_reference_model(sd)
    fp = state_fingerprint(sd)
    with torch.no_grad():
        # stage A: ragged lengths incl. a non-multiple of the hop and the minimum useful size
        for tag, n in (("a_n4410", 4410), ("a_n30001", 30001)):
            wav = O.synth_clips(3, n, seed=11)
            sp, cos, sin = model.f_helper.wav_to_spectrogram_phase(wav[:, None, :])
            mel = model.mel(sp.permute(0, 1, 3, 2)).permute(0, 1, 3, 2)
            np.savez_compressed(os.path.join(GOLD, f"stage_{tag}.npz"), wav=wav.numpy(), sp=sp.numpy(),
                                mel=mel.numpy(), fingerprint=fp)
        # stage B: the reference's own smoke shape (unet.py:107, T=101) and a 10 s clip (T=1001)
        for tag, t, b in (("b_t101", 101, 2), ("b_t1001", 1001, 1)):
            g = torch.Generator().manual_seed(5 + t)
            mel_orig = 10 ** (torch.randn(b, 1, t, 128, generator=g) * 0.8 - 1.0)
            mel_orig[:, :, ::17, ::5] = 0.0                      # exercise the 1e-8 clip of to_log
            out = model(mel_orig)["mel"]
            np.savez_compressed(os.path.join(GOLD, f"stage_{tag}.npz"), mel_orig=mel_orig.numpy(),
                                log_mel=out.numpy(), fingerprint=fp)
        # end to end: 1 s clips (B=2) and one 10 s clip through the handler body
        for tag, n, b in (("e2e_1s", 44100, 2), ("e2e_10s", 441000, 1)):
            wav = O.synth_clips(b, n, seed=21 + b)
            col = {}
            out = ref_import.reference_handler_batch(model, wav, collect=col)
            np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), wav=wav.numpy(), out=out.numpy(),
                                log_mel=torch.cat(col["log_mel"]).numpy(), fingerprint=fp)
            print(tag, "out rms", float(out.pow(2).mean().sqrt()))
    print("golden written to", GOLD)


def main_ssr():
    """SURVEY.md 8(f) row 1 (next path): models/components/unet_v2.py UNetResComplex_100Mb imported unmodified, with
    the UNet tensors of make_state(SEED) under the SSR prefix (same shapes, models/ssr_unet.py:49).  Its STFT/ISTFT
    helpers are the oracle's restatements (torchlibrosa is absent)."""
    from voicefixer_main_b200.arch import UNET_PREFIX
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_state(SEED)
    ssr = {k.replace(UNET_PREFIX, "generator.unet."): v for k, v in sd.items() if k.startswith(UNET_PREFIX)}
    net = ref_import.build_reference_unet_v2(ssr)
    fp = state_fingerprint(sd)
    with torch.no_grad():
        wav = O.synth_clips(2, 63 * 441 + 200, seed=51)[:, None, :]     # T = 64 frames, ragged sample count
        sp, cos, sin = net.f_helper.wav_to_spectrogram_phase(wav)
        out = net(sp, wav)["wav"]
        mag = O.unet_v2_forward(ssr, sp)                                  # bit-identical to the module's out_mag (tested)
    np.savez_compressed(os.path.join(GOLD, "ssr_t64.npz"), wav=wav[:, 0].numpy(), out_mag=mag.numpy(),
                        out=out[:, 0].numpy(), fingerprint=fp)
    print("ssr_t64 out rms", float(out.pow(2).mean().sqrt()))


def main_small():
    """the `unet_small` analysis module (models/components/unet_small.py) imported unmodified,
    through Generator.forward's arithmetic (gsr_voicefixer.py:86-91: unet(to_log(mel)) + to_log(mel))."""
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_state(SEED)
    net = ref_import.build_reference_unet_small(sd)
    from tools.pytorch.pytorch_util import to_log
    g = torch.Generator().manual_seed(77)
    mel_orig = 10 ** (torch.randn(1, 1, 101, 128, generator=g) * 0.8 - 1.0)
    with torch.no_grad():
        out = net(to_log(mel_orig))["mel"] + to_log(mel_orig)
    np.savez_compressed(os.path.join(GOLD, "stage_b_small_t101.npz"), mel_orig=mel_orig.numpy(), log_mel=out.numpy(),
                        fingerprint=state_fingerprint(sd))
    print("stage_b_small_t101 written")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ssr":
        main_ssr()
    elif len(sys.argv) > 1 and sys.argv[1] == "small":
        main_small()
    else:
        main()
