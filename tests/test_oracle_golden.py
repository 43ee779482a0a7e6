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


def _ssr_state(state):
    from voicefixer_main_b200.arch import UNET_PREFIX
    return {k.replace(UNET_PREFIX, "generator.unet."): v for k, v in state.items() if k.startswith(UNET_PREFIX)}


def test_istft_restatement_properties():
    """The ISTFT lives in torchlibrosa (absent): the restatement is pinned by definition (torch.istft computes the same
    windowed overlap-add with squared-window normalisation) and by the STFT round trip."""
    x = O.synth_clips(2, 30011, seed=13)
    real, imag = O.stft_conv_dft(x)
    y = O.istft(real, imag, x.shape[1])
    assert y.shape == x.shape and y.dtype == torch.float32
    assert float((y - x).abs().max()) < 1e-5
    spec = torch.complex(real[:, 0].double(), imag[:, 0].double()).transpose(1, 2)
    ref = torch.istft(spec, n_fft=O.N_FFT, hop_length=O.HOP, win_length=O.N_FFT, window=O.hann_periodic(), center=True,
                      length=x.shape[1])
    assert float((y.double() - ref).abs().max()) < 1e-6
    # a shorter request keeps the head (torchlibrosa _trim_edges), linearity in the spectrum
    assert torch.equal(O.istft(real, imag, 1000), y[:, :1000])
    y2 = O.istft(2 * real, 2 * imag, x.shape[1])
    assert float((y2 - 2 * y).abs().max()) < 1e-6


def test_ssr_path_against_reference_golden(state, golden_fingerprint_ok):
    """SURVEY.md 8(f) row 1 (next path): unet_v2 magnitude branch + input phase + ISTFT vs the reference module's output."""
    g = load_golden("ssr_t64.npz")
    ssr = _ssr_state(state)
    wav = torch.from_numpy(g["wav"])[:, None, :]
    with torch.no_grad():
        sp, _, _ = O.wav_to_spectrogram_phase(wav)
        mag = O.unet_v2_forward(ssr, sp)
        out = O.ssr_forward(ssr, wav)
    assert mag.shape == g["out_mag"].shape == (2, 1, 64, 1025)
    assert float((mag - torch.from_numpy(g["out_mag"])).abs().max()) < 2e-5 * float(np.abs(g["out_mag"]).max())
    assert bool((mag[..., 1024] == 0).all())                       # padded last bin (unet_v2.py:128)
    assert out.shape == (2, 1, wav.shape[2])
    assert float((out[:, 0] - torch.from_numpy(g["out"])).abs().max()) < 2e-5
