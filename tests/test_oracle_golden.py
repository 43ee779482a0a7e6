"""Oracle (oracle/vf_oracle.py) against the golden vectors generated from the reference
itself (oracle/make_golden.py).  Runs anywhere (no GPU, no /root/reference)."""
import numpy as np
import torch

from conftest import load_golden
from oracle import vf_oracle as O


def test_stage_a_against_reference_golden(golden_fingerprint_ok):
    for name in ("stage_a_n4410.npz", "stage_a_n30001.npz"):
        g = load_golden(name)
        wav = torch.from_numpy(g["wav"])
        sp, mel = O.pre(wav[:, None, :])
        assert sp.shape == g["sp"].shape and mel.shape == g["mel"].shape
        assert sp.shape[2] == 1 + wav.shape[1] // 441
        np.testing.assert_allclose(sp.numpy(), g["sp"], rtol=0, atol=1e-6 * float(g["sp"].max()))
        np.testing.assert_allclose(mel.numpy(), g["mel"], rtol=0, atol=1e-6 * float(g["mel"].max()))
        # the fp64 FFT form agrees with the reference's fp32 conv-DFT to fp32 rounding
        sp64, _ = O.pre(wav[:, None, :], exact=True)
        assert float((sp64 - torch.from_numpy(g["sp"])).abs().max()) < 3e-6 * float(g["sp"].max())


def test_stage_b_against_reference_golden(state, golden_fingerprint_ok):
    g = load_golden("stage_b_t101.npz")
    with torch.no_grad():
        out = O.generator_forward(state, torch.from_numpy(g["mel_orig"]))
    err = float((out - torch.from_numpy(g["log_mel"])).abs().max())
    assert err < 2e-5, err
    # last mel bin is a pure pass-through of the input (unet.py:78,99)
    lm = O.to_log(torch.from_numpy(g["mel_orig"]))
    assert torch.equal(out[..., 127], lm[..., 127])


def test_end_to_end_against_reference_golden(state, golden_fingerprint_ok):
    g = load_golden("e2e_1s.npz")
    st = {}
    with torch.no_grad():
        out = O.restore(state, torch.from_numpy(g["wav"]), stages=st)
    assert out.shape == g["out"].shape
    assert float((torch.cat(st["log_mel"]) - torch.from_numpy(g["log_mel"])).abs().max()) < 2e-5
    rms_err = float((out - torch.from_numpy(g["out"])).pow(2).mean().sqrt())
    assert rms_err < 1e-6, rms_err


def test_segment_loop_and_trim():
    # handler semantics on a ragged multi-segment input with a tiny "segment" length
    est = torch.arange(20.)[None, None]
    assert O.trim_center(est, 14).tolist() == [[list(map(float, range(3, 17)))]]
    assert O.trim_center(est, 20) is est
    assert O.to_int16(np.array([0.5, -0.5, 0.99999], dtype=np.float32)).tolist() == [16384, -16384, 32767]
    x = torch.tensor([[[0.5, -2.0]], [[0.25, 0.5]]])
    y = O.peak_normalize(x)
    assert y[0].tolist() == [[0.25, -1.0]] and torch.equal(y[1], x[1])
