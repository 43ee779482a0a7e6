"""CPU oracle for the VoiceFixer inference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this file; the product (voicefixer_main_b200/) never
does and fails loudly when its CUDA library is missing.

Every function restates, in plain fp32 PyTorch on the CPU, one step of the
reference's handler (eval_gsr_voicefixer.py:37-77) and cites the lines it
follows.  Pinning status:

* stage B (analysis ResUNet), mel filterbank, to_log/from_log, trim_center:
  PINNED - tests/test_oracle_vs_reference.py imports the reference's own modules
  unmodified (oracle/ref_import.py) in the build container and requires
  bit-level / 1e-6 agreement; tests/golden/*.npz hold outputs generated from
  the reference itself (oracle/make_golden.py) for the GPU box, where
  /root/reference does not exist.
* stage A (STFT): the arithmetic lives in torchlibrosa==0.0.7
  (requirements.txt:10; constructed at tools/pytorch/modules/fDomainHelper.py:26-28),
  which is absent.  `stft_conv_dft` restates its published algorithm (reflect pad,
  periodic-hann-windowed DFT matrix applied as a strided conv1d); pinned by
  construction against an fp64 FFT (`stft_exact`), not by any reference test
  (the reference has none).
* stage C (vocoder): lives in the unpinned `voicefixer` PyPI package
  (requirements.txt:6; call site eval_gsr_voicefixer.py:66), source and weights
  absent -> PARITY UNPINNED.  `vocoder_forward` restates the published generator
  design under voicefixer_main_b200.arch.VocoderConfig.
* next path (SSR): `unet_v2_forward` is PINNED (bit-identical to
  models/components/unet_v2.py imported unmodified, golden ssr_t64.npz); `istft` restates
  torchlibrosa's ISTFT (absent) and is pinned by torch.istft and the STFT round trip only.
"""
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from voicefixer_main_b200.arch import (BN_EPS, DEC_CHANNELS, ENC_CHANNELS, LRELU_SLOPE,
                                       UNET_PREFIX, VocoderConfig, padded_frames)

N_FFT = 2048          # config/vctk_base_voicefixer_unet.json:72
HOP = 441             # config/vctk_base_voicefixer_unet.json:73
N_MELS = 128          # config/vctk_base_voicefixer_unet.json:71
SR = 44100            # config/vctk_base_voicefixer_unet.json:68
SEG_SAMPLES = SR * 60  # eval_gsr_voicefixer.py:47


# ----------------------------------------------------------------------------
# Stage A: STFT magnitude + mel projection
# ----------------------------------------------------------------------------
def hann_periodic(n: int = N_FFT, dtype=torch.float64) -> torch.Tensor:
    """librosa.filters.get_window('hann', n, fftbins=True) as torchlibrosa builds it."""
    k = torch.arange(n, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)).to(dtype)


def reflect_pad(x: torch.Tensor, pad: int = N_FFT // 2) -> torch.Tensor:
    """center=True, pad_mode='reflect' (fDomainHelper.py:26-28 -> torchlibrosa STFT.forward)."""
    return F.pad(x[:, None, :], (pad, pad), mode="reflect")[:, 0, :]


_DFT_CACHE = {}


def _dft_weights():
    if "w" not in _DFT_CACHE:
        n = torch.arange(N_FFT, dtype=torch.float64)[:, None]
        k = torch.arange(N_FFT // 2 + 1, dtype=torch.float64)[None, :]
        ang = 2.0 * math.pi * n * k / N_FFT
        win = hann_periodic()[:, None]
        wr = (torch.cos(ang) * win).T.float()[:, None, :]     # [1025,1,2048]
        wi = (-torch.sin(ang) * win).T.float()[:, None, :]
        _DFT_CACHE["w"] = (wr, wi)
    return _DFT_CACHE["w"]


def stft_conv_dft(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Literal restatement of torchlibrosa 0.0.7 STFT.forward: two fp32 conv1d
    with the windowed DFT matrix, stride = hop.  x [B,N] -> real, imag [B,1,T,1025]."""
    wr, wi = _dft_weights()
    xp = reflect_pad(x.float())[:, None, :]
    real = F.conv1d(xp, wr, stride=HOP)
    imag = F.conv1d(xp, wi, stride=HOP)
    return real[:, None].transpose(2, 3).contiguous(), imag[:, None].transpose(2, 3).contiguous()


def stft_exact(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same transform in fp64 via FFT: the value both the reference's fp32 conv-DFT
    and the CUDA fp32 FFT approximate.  Returns fp64 real, imag [B,1,T,1025]."""
    xp = reflect_pad(x.double())
    frames = xp.unfold(-1, N_FFT, HOP) * hann_periodic()      # [B,T,2048]
    spec = torch.fft.rfft(frames, dim=-1)
    return spec.real[:, None], spec.imag[:, None]


def spectrogram_phase(real, imag, eps: float = 1e-8):
    """fDomainHelper.py:60-65: mag = clamp(r^2+i^2, eps, inf)^0.5, cos = r/mag, sin = i/mag."""
    mag = torch.clamp(real ** 2 + imag ** 2, eps, np.inf) ** 0.5
    return mag, real / mag, imag / mag


def wav_to_spectrogram_phase(x: torch.Tensor, exact: bool = False):
    """fDomainHelper.py:67-89 for x [B,C,N] -> mag, cos, sin each [B,C,T,1025]."""
    outs = [[], [], []]
    for c in range(x.shape[1]):
        real, imag = (stft_exact if exact else stft_conv_dft)(x[:, c, :])
        for lst, t in zip(outs, spectrogram_phase(real, imag)):
            lst.append(t)
    return tuple(torch.cat(l, dim=1) for l in outs)


def _hz_to_mel(f: float) -> float:
    return 2595.0 * math.log10(1.0 + f / 700.0)               # mel_scale.py:66-97 (htk)


def mel_filterbank(n_freqs: int = N_FFT // 2 + 1, n_mels: int = N_MELS, sr: int = SR) -> torch.Tensor:
    """tools/pytorch/mel_scale.py:131-221 with f_min=0, f_max=sr//2, norm=None, htk:
    triangular filters, fb [n_freqs, n_mels] fp32, computed with the same fp32 torch
    ops in the same order so the result is bit-identical to the reference buffer."""
    all_freqs = torch.linspace(0, sr // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel(0.0), _hz_to_mel(float(sr // 2)), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)           # mel_scale.py:99-129
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def mel_project(sp: torch.Tensor, fb: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gsr_voicefixer.py:180: mel(sp.permute(0,1,3,2)).permute(0,1,3,2) == sp @ fb."""
    fb = mel_filterbank() if fb is None else fb
    return torch.matmul(sp, fb.to(sp.dtype))


def pre(wav: torch.Tensor, exact: bool = False):
    """VoiceFixer.pre (gsr_voicefixer.py:178-181): wav [B,1,N] -> sp [B,1,T,1025], mel [B,1,T,128]."""
    sp, _, _ = wav_to_spectrogram_phase(wav, exact=exact)
    return sp, mel_project(sp)


def to_log(x: torch.Tensor) -> torch.Tensor:
    """tools/pytorch/pytorch_util.py:157-159."""
    assert torch.sum(x < 0) == 0, "input has negative values"
    return torch.log10(torch.clip(x, min=1e-8))


def from_log(x: torch.Tensor) -> torch.Tensor:
    """tools/pytorch/pytorch_util.py:161-163."""
    return 10 ** torch.clip(x, min=-np.inf, max=5)


# ----------------------------------------------------------------------------
# Stage B: analysis ResUNet (models/components/unet.py, modules.py)
# ----------------------------------------------------------------------------
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _conv_block_res(x, sd, p):
    """modules.py:263-271."""
    origin = x
    x = F.conv2d(F.leaky_relu(_bn(x, sd, p + ".bn1"), LRELU_SLOPE), sd[p + ".conv1.weight"], padding=1)
    x = F.conv2d(F.leaky_relu(_bn(x, sd, p + ".bn2"), LRELU_SLOPE), sd[p + ".conv2.weight"], padding=1)
    if (p + ".shortcut.weight") in sd:
        return F.conv2d(origin, sd[p + ".shortcut.weight"], sd[p + ".shortcut.bias"]) + x
    return origin + x


def unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = UNET_PREFIX) -> torch.Tensor:
    """UNetResComplex_100Mb.forward (unet.py:60-103): x [B,1,T,128] -> [B,1,T,128]."""
    sd = {k[len(prefix):]: v.to(x.dtype) if v.is_floating_point() else v
          for k, v in sd.items() if k.startswith(prefix)}
    origin_len = x.shape[2]
    x = F.pad(x, (0, 0, 0, padded_frames(origin_len) - origin_len))    # unet.py:75-77
    x = x[..., 0:x.shape[-1] - 1]                                       # unet.py:78
    skips = []
    for i in range(1, len(ENC_CHANNELS) + 1):                           # modules.py:177-184
        for j in range(1, 5):
            x = _conv_block_res(x, sd, f"encoder_block{i}.conv_block{j}")
        skips.append(x)
        x = F.avg_pool2d(x, kernel_size=(2, 2))
    x = _conv_block_res(x, sd, "conv_block7")
    for i in range(1, len(DEC_CHANNELS) + 1):                           # modules.py:212-220
        p = f"decoder_block{i}"
        x = F.conv_transpose2d(F.relu(_bn(x, sd, p + ".bn1")), sd[p + ".conv1.weight"], stride=2)
        x = x[:, :, 0:-1, :]                                            # prune, both=False
        x = torch.cat((x, skips[-i]), dim=1)
        for j in range(2, 6):
            x = _conv_block_res(x, sd, f"{p}.conv_block{j}")
    x = _conv_block_res(x, sd, "after_conv_block1")
    x = F.conv2d(x, sd["after_conv2.weight"], sd["after_conv2.bias"])
    x = F.pad(x, (0, 1))                                                # unet.py:99
    return x[:, :, 0:origin_len, :]


def unet_v2_forward(sd: Dict[str, torch.Tensor], sp: torch.Tensor, prefix: str = "generator.unet.") -> torch.Tensor:
    """SURVEY.md 8(f) row 1 (next path, BASELINE config 3): the magnitude branch of
    models/components/unet_v2.py:86-132.  sp [B,1,T,1025] linear magnitude -> out_mag [B,1,T,1025].
    Same blocks as unet_forward, but on F = 1024 bins, the decoders prune BOTH the last time row and the last
    frequency column (`both=True`, modules.py:203-210) and the output is used as the magnitude itself (:134-136)."""
    sd = {k[len(prefix):]: v.to(sp.dtype) if v.is_floating_point() else v
          for k, v in sd.items() if k.startswith(prefix)}
    x = sp
    origin_len = x.shape[2]
    x = F.pad(x, (0, 0, 0, padded_frames(origin_len) - origin_len))    # unet_v2.py:101-104
    x = x[..., 0:x.shape[-1] - 1]                                       # unet_v2.py:108
    skips = []
    for i in range(1, len(ENC_CHANNELS) + 1):
        for j in range(1, 5):
            x = _conv_block_res(x, sd, f"encoder_block{i}.conv_block{j}")
        skips.append(x)
        x = F.avg_pool2d(x, kernel_size=(2, 2))
    x = _conv_block_res(x, sd, "conv_block7")
    for i in range(1, len(DEC_CHANNELS) + 1):
        p = f"decoder_block{i}"
        x = F.conv_transpose2d(F.relu(_bn(x, sd, p + ".bn1")), sd[p + ".conv1.weight"], stride=2)
        x = x[:, :, 0:-1, 0:-1]                                         # prune, both=True (modules.py:207-208)
        x = torch.cat((x, skips[-i]), dim=1)
        for j in range(2, 6):
            x = _conv_block_res(x, sd, f"{p}.conv_block{j}")
    x = _conv_block_res(x, sd, "after_conv_block1")
    x = F.conv2d(x, sd["after_conv2.weight"], sd["after_conv2.bias"])
    x = F.pad(x, (0, 1))                                                # unet_v2.py:128
    return x[:, :, 0:origin_len, :][:, 0:1]                             # unet_v2.py:129-131


def istft(real: torch.Tensor, imag: torch.Tensor, length: int) -> torch.Tensor:
    """FDomainHelper.istft (fDomainHelper.py:30-32, 127) = torchlibrosa 0.0.7 ISTFT.forward with n_fft = win = 2048,
    hop = 441, periodic hann, center=True.  The package is absent from this image: restated from its published
    algorithm [recollection] - mirror the half spectrum, inverse DFT of every frame, multiply by the window,
    overlap-add, divide by the overlap-added squared window (clamped at 1e-11), drop n_fft/2 samples and keep
    `length`.  PINNED ONLY BY PROPERTY: tests check it against torch.istft (same definition) and the
    stft -> istft round trip.  real, imag [B,1,T,1025] -> [B,length]."""
    spec = torch.complex(real[:, 0].double(), imag[:, 0].double())       # [B,T,1025]
    frames = torch.fft.irfft(spec, n=N_FFT, dim=-1)                      # = Re(idft of the mirrored full spectrum) / n_fft
    win = hann_periodic()
    frames = frames * win
    b, t, _ = frames.shape
    out_len = (t - 1) * HOP + N_FFT
    y = F.fold(frames.transpose(1, 2), output_size=(1, out_len), kernel_size=(1, N_FFT), stride=(1, HOP))[:, 0, 0, :]
    wsum = F.fold((win ** 2)[None, :, None].repeat(1, 1, t), output_size=(1, out_len), kernel_size=(1, N_FFT),
                  stride=(1, HOP))[0, 0, 0, :]
    y = y / torch.clamp(wsum, 1e-11, np.inf)
    return y[:, N_FFT // 2:N_FFT // 2 + length].to(real.dtype)


def ssr_forward(sd, wav: torch.Tensor, exact_stft: bool = False) -> torch.Tensor:
    """SSR_UNet.forward (models/ssr_unet.py:144-155) after pre (:140-143): wav [B,1,N] -> restored wav [B,1,N]:
    magnitude from the UNet, phase of the input (unet_v2.py:97,138-139), ISTFT to the input length (:141-143)."""
    sp, cos_in, sin_in = wav_to_spectrogram_phase(wav, exact=exact_stft)
    out_mag = unet_v2_forward(sd, sp.float())
    y = istft(out_mag * cos_in.float(), out_mag * sin_in.float(), wav.shape[2])
    return y[:, None, :]


def generator_forward(sd, mel_orig: torch.Tensor) -> torch.Tensor:
    """Generator.forward (gsr_voicefixer.py:86-91): returns the log10 mel."""
    logm = to_log(mel_orig)
    return unet_forward(sd, logm) + logm


def amp_to_original_f(mel_est, mel_target, cutoff: float = 0.2):
    """tools/utils.py:50-55."""
    fd = mel_target.size()[-1]
    e_est = torch.mean(mel_est[..., 5:int(fd * cutoff)], dim=(2, 3))
    e_tgt = torch.mean(mel_target[..., 5:int(fd * cutoff)], dim=(2, 3))
    return mel_est * (e_tgt / e_est)[..., None, None], mel_target


# ----------------------------------------------------------------------------
# Stage C: vocoder restatement (PARITY UNPINNED, see header)
# ----------------------------------------------------------------------------
def mel_weight(cfg: VocoderConfig) -> torch.Tensor:
    x = torch.arange(cfg.num_mels, dtype=torch.float64)
    return (cfg.mel_weight_a * torch.exp(cfg.mel_weight_b * x)).float()


def vocoder_condition(mel: torch.Tensor, cfg: VocoderConfig) -> torch.Tensor:
    """Vocoder.forward prologue: mel [B,1,T,128] linear -> conditions [B,128,T+pad]."""
    mel = mel / mel_weight(cfg).to(mel.dtype)
    s = 20.0 * torch.log10(torch.clamp(torch.abs(mel), min=cfg.amp_floor)) - cfg.ref_db
    s = torch.clip((s - cfg.min_db) / (-cfg.min_db), 0, 1)
    c = s[:, 0].transpose(1, 2)
    pad_tail = c.size(-1) % 2 + cfg.tail_pad_base
    tail = torch.zeros([c.size(0), cfg.num_mels, pad_tail], dtype=c.dtype) + cfg.tail_pad_value
    return torch.cat([c, tail], dim=-1)


def vocoder_generator(sd, c: torch.Tensor, cfg: VocoderConfig, prefix: str = "vocoder.") -> torch.Tensor:
    g = lambda k: sd[prefix + k].to(c.dtype)
    x = c
    for i in range(cfg.cond_layers):
        x = F.elu(F.conv1d(x, g(f"condnet.{i}.weight"), g(f"condnet.{i}.bias"), padding=1))
    hk = cfg.stem_kernel // 2
    x = F.conv1d(F.pad(x, (hk, hk), mode="reflect"), g("stem.weight"), g("stem.bias"))
    x = F.leaky_relu(x, cfg.stage_slope)
    for s, (scale, depth) in enumerate(zip(cfg.upsample_scales, cfg.resstack_depth)):
        x = F.conv_transpose1d(x, g(f"up.{s}.weight"), g(f"up.{s}.bias"), stride=scale,
                               padding=scale // 2 + scale % 2, output_padding=scale % 2)
        for i in range(depth):
            d = cfg.dilation(i)
            pk = (cfg.resstack_kernel * d - d) // 2
            h = F.conv1d(F.leaky_relu(x, cfg.res_slope), g(f"res.{s}.{i}.a.weight"), g(f"res.{s}.{i}.a.bias"),
                         dilation=d, padding=pk)
            h = F.conv1d(F.leaky_relu(h, cfg.res_slope), g(f"res.{s}.{i}.b.weight"), g(f"res.{s}.{i}.b.bias"),
                         padding=cfg.resstack_kernel // 2)
            x = x + h
        x = F.leaky_relu(x, cfg.stage_slope)
    x = F.conv1d(F.pad(x, (hk, hk), mode="reflect"), g("tail.weight"), g("tail.bias"))
    return torch.tanh(x) if getattr(cfg, "tail_tanh", True) else x


def vocoder_forward(sd, mel: torch.Tensor, cfg: Optional[VocoderConfig] = None) -> torch.Tensor:
    """mel [B,1,T,128] linear -> wav [B,1,(T + T%2 + 4) * 441]."""
    cfg = cfg or VocoderConfig()
    return vocoder_generator(sd, vocoder_condition(mel, cfg), cfg)


# ----------------------------------------------------------------------------
# Handler tail: peak normalise, trim, int16
# ----------------------------------------------------------------------------
def peak_normalize(out: torch.Tensor) -> torch.Tensor:
    """eval_gsr_voicefixer.py:68-70, applied per clip (the reference only ever sees batch 1)."""
    peak = out.abs().amax(dim=(1, 2), keepdim=True)
    return torch.where(peak > 1.0, out / peak, out)


def trim_center(est: torch.Tensor, ref_len: int) -> torch.Tensor:
    """tools/utils.py:57-70, est longer than or equal to ref (the vocoder case)."""
    diff = abs(est.shape[-1] - ref_len)
    if est.shape[-1] == ref_len:
        return est
    assert est.shape[-1] > ref_len
    est = est[..., int(diff // 2):-int(diff // 2)]
    return est[..., :ref_len]


def to_int16(frames: np.ndarray) -> np.ndarray:
    """tools/file/wav.py:22-24: scale by 2^15 and truncate toward zero (astype(np.short))."""
    frames = frames.astype(np.float32) * np.float32(2 ** 15)
    return frames.astype(np.short)


def restore(sd, wav: torch.Tensor, cfg: Optional[VocoderConfig] = None, exact_stft: bool = False,
            seg_samples: int = SEG_SAMPLES, stages: Optional[dict] = None, unify_energy: bool = False) -> torch.Tensor:
    """handler() of eval_gsr_voicefixer.py:37-77 for a batch of equal-length clips
    wav [B,N] -> [B,N]: independent 60 s segments, stages A->B->C, peak normalise, trim, concat."""
    cfg = cfg or VocoderConfig()
    res = []
    n = wav.shape[-1]
    bp = seg_samples
    while bp < n + seg_samples:                                  # eval_gsr_voicefixer.py:49
        seg = wav[:, bp - seg_samples:bp]
        _, mel_noisy = pre(seg[:, None, :], exact=exact_stft)
        log_mel = generator_forward(sd, mel_noisy.float())
        denoised = from_log(log_mel)
        if unify_energy:                                          # eval_gsr_voicefixer.py:54-55
            denoised, _ = amp_to_original_f(mel_est=denoised, mel_target=mel_noisy.float())
        out = vocoder_forward(sd, denoised, cfg)
        out = peak_normalize(out)
        out = trim_center(out, seg.shape[-1])
        if stages is not None:
            stages.setdefault("mel_noisy", []).append(mel_noisy)
            stages.setdefault("log_mel", []).append(log_mel)
        res.append(out)
        bp += seg_samples
    return torch.cat(res, -1)[:, 0, :]


def synth_clips(batch: int, n_samples: int, seed: int = 1234, kind: str = "speech") -> torch.Tensor:
    """SURVEY.md 8(d) synthetic inputs: 0.1*randn, or harmonic 'speech-like' clips."""
    g = torch.Generator().manual_seed(seed)
    if kind == "noise":
        return 0.1 * torch.randn(batch, n_samples, generator=g)
    t = torch.arange(n_samples, dtype=torch.float64) / SR
    out = []
    for _ in range(batch):
        f0 = 80 + 220 * torch.rand(1, generator=g).item()
        sig = torch.zeros(n_samples, dtype=torch.float64)
        for h in range(1, 25):
            amp = torch.rand(1, generator=g).item() / h
            ph = 2 * math.pi * torch.rand(1, generator=g).item()
            vib = 1.0 + 0.01 * torch.sin(2 * math.pi * 5.0 * t + ph)
            sig += amp * torch.sin(2 * math.pi * f0 * h * t * vib + ph)
        env = 0.55 + 0.45 * torch.sin(2 * math.pi * (0.5 + torch.rand(1, generator=g).item()) * t)
        sig = sig * env + 0.003 * torch.randn(n_samples, generator=g, dtype=torch.float64)
        sig = sig / sig.abs().max() * (0.3 + 0.7 * torch.rand(1, generator=g).item())
        out.append(sig.float())
    return torch.stack(out)
