"""Pin the restatement in oracle/vf_oracle.py against the reference's own modules,
imported unmodified (build container only; skipped where /root/reference is absent)."""
import pytest
import torch

from oracle import ref_import, vf_oracle as O

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_model(state):
    model, _ = ref_import.build_reference_model(state)
    return model


def test_mel_filterbank_bit_identical(ref_model):
    fb = O.mel_filterbank()
    assert torch.equal(ref_model.mel.fb, fb)
    assert int((fb != 0).sum()) == 2018            # SURVEY.md 8(a) a5 probe


def test_log_helpers_match_reference():
    ref_import.install_shims()
    from tools.pytorch.pytorch_util import from_log, to_log
    x = torch.rand(3, 1, 7, 128) * 3
    x[0, 0, 0, :4] = 0
    assert torch.equal(to_log(x), O.to_log(x))
    y = torch.randn(3, 1, 7, 128) * 4
    assert torch.equal(from_log(y), O.from_log(y))
    with pytest.raises(AssertionError):
        O.to_log(-x - 1)
    with pytest.raises(AssertionError):
        to_log(-x - 1)


def test_unet_restatement_matches_reference(ref_model, state):
    g = torch.Generator().manual_seed(3)
    for t in (64, 101, 130):                          # multiple of 64, the reference smoke shape, ragged
        mel = 10 ** (torch.randn(2, 1, t, 128, generator=g) - 1)
        with torch.no_grad():
            ref = ref_model(mel)["mel"]
            mine = O.generator_forward(state, mel)
        assert ref.shape == mine.shape == (2, 1, t, 128)
        assert float((ref - mine).abs().max()) < 2e-5


def test_handler_restatement_matches_reference(ref_model, state):
    wav = O.synth_clips(2, 30000, seed=9)
    with torch.no_grad():
        ref = ref_import.reference_handler_batch(ref_model, wav, seg_samples=12000)   # 3 segments, last ragged
        mine = O.restore(state, wav, seg_samples=12000)
    assert ref.shape == mine.shape == wav.shape
    assert float((ref - mine).abs().max()) < 1e-5


def test_trim_center_matches_reference():
    ref_import.install_shims()
    from tools.utils import trim_center
    for le, lr in ((20, 14), (443646, 441000), (16, 16)):
        est = torch.arange(float(le))[None, None]
        ref = torch.zeros(1, 1, lr)
        assert torch.equal(trim_center(est, ref)[0], O.trim_center(est, lr))


def test_unet_v2_ssr_restatement_matches_reference(state):
    """Next path: models/components/unet_v2.py imported unmodified vs the restatement."""
    from voicefixer_main_b200.arch import UNET_PREFIX
    ssr = {k.replace(UNET_PREFIX, "generator.unet."): v for k, v in state.items() if k.startswith(UNET_PREFIX)}
    net = ref_import.build_reference_unet_v2(ssr)
    for n in (63 * 441, 70 * 441 + 17):                   # T = 64 (no time padding) and T = 71 -> T' = 128
        wav = O.synth_clips(1, n, seed=n)[:, None, :]
        with torch.no_grad():
            sp, _, _ = O.wav_to_spectrogram_phase(wav)
            ref = net(sp, wav)["wav"]
            mine = O.ssr_forward(ssr, wav)
        assert ref.shape == mine.shape == (1, 1, n)
        assert float((ref - mine).abs().max()) < 1e-5


def test_unet_small_is_the_same_network(ref_model, state):
    """models/components/unet_small.py imported unmodified.  In this reference its *Res1B blocks
    hold four ConvBlockRes each (modules.py:112-165), i.e. the layers and keys of unet.py: the product maps
    `unet_small: true` onto the same plan, which this test justifies bit for bit."""
    small = ref_import.build_reference_unet_small(state)
    g = torch.Generator().manual_seed(4)
    mel = 10 ** (torch.randn(1, 1, 101, 128, generator=g) - 1)
    with torch.no_grad():
        a = small(O.to_log(mel))["mel"] + O.to_log(mel)
        b = ref_model(mel)["mel"]
        c = O.generator_forward(state, mel)
    assert torch.equal(a, b)
    assert float((a - c).abs().max()) < 2e-5
