"""I/O edges of handler() on the GPU: resampling to the model rate and the mel metrics.

* `resample_poly` / `resample_to`: zero-phase polyphase FIR rate conversion with the arithmetic of
  scipy.signal.resample_poly (Kaiser beta = 5 low-pass of 20 * max(up, down) + 1 taps), which is what the reference
  itself uses to change rates (tools/dsp/lowpass.py:138-141).  The reference's `load_wav` (tools/utils.py:46-48) calls
  librosa.load(sr=44100), whose resampler depends on the installed librosa (soxr_hq / kaiser_best) and is not available
  offline; the FIR here is of the same class (windowed-sinc, > 60 dB stop band) and is checked against scipy.
* `AudioMetrics.lsd` / `.sispec`: evaluation_proc/metrics.py:83-95 for [B, C, T, F] tensors on the device, the two
  metrics handler() logs per segment when a target is given (eval_gsr_voicefixer.py:56-64).  `ssim` is CPU numpy in
  the reference (skimage, metrics.py:97-106) and is not part of this path.

The filter design is host arithmetic (numpy); every sample / reduction is computed by libb200vf kernels (edges.cu).
"""
import ctypes
from fractions import Fraction

import numpy as np
import torch

from . import _lib as L
from .model import Engine, _check_in, _ptr, _stream


def design_filter(up: int, down: int) -> np.ndarray:
    """scipy.signal.resample_poly's default FIR: firwin(2 * half + 1, 1 / max(up, down), window=('kaiser', 5.0)) * up with
    half = 10 * max(up, down) (scipy/signal/_signaltools.py), restated with numpy: ideal low-pass sinc x Kaiser window,
    unity DC gain.  float64 taps, returned as float32."""
    max_rate = max(up, down)
    half = 10 * max_rate
    n = np.arange(-half, half + 1, dtype=np.float64)
    fc = 1.0 / max_rate                                    # cutoff as a fraction of Nyquist
    h = fc * np.sinc(fc * n) * np.kaiser(2 * half + 1, 5.0)
    h /= h.sum()
    return (h * up).astype(np.float32)


def resample_poly(eng: Engine, x: torch.Tensor, up: int, down: int) -> torch.Tensor:
    """x [B, N] float32 on the engine's device -> [B, ceil(N * up / down)]."""
    x = _check_in(x, eng.device, "x")
    g = np.gcd(int(up), int(down))
    up, down = int(up) // g, int(down) // g
    if up == down == 1:
        return x.clone()
    b, n = x.shape
    n_out = (n * up + down - 1) // down
    taps = torch.from_numpy(design_filter(up, down)).to(eng.device)
    out = torch.empty(b, n_out, device=eng.device)
    with torch.cuda.device(eng.device):
        eng._ck(eng.lib.vf_resample_poly(eng.ctx, _ptr(x), b, n, up, down, _ptr(taps), taps.numel(), _ptr(out), n_out, _stream()))
    return out


def resample_to(eng: Engine, x: torch.Tensor, rate_in: int, rate_out: int = 44100) -> torch.Tensor:
    """load_wav's rate conversion (tools/utils.py:46-48: librosa.load(path, sr=44100)) for a decoded signal."""
    fr = Fraction(int(rate_out), int(rate_in))
    return resample_poly(eng, x, fr.numerator, fr.denominator)


class AudioMetrics:
    """evaluation_proc/metrics.py:25-106, the parts handler() calls on device tensors: lsd and sispec."""

    def __init__(self, owner, rate: int = 44100):
        self.rate = rate
        self._owner = owner

    def _eng(self) -> Engine:
        return self._owner._engine() if hasattr(self._owner, "_engine") else self._owner

    def lsd(self, est: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """metrics.py:83-87 (non-log inputs [B, C, T, F]) -> [B, C, 1, 1]."""
        eng = self._eng()
        est, target = _check_in(est, eng.device, "est"), _check_in(target, eng.device, "target")
        assert est.dim() == 4 and est.shape == target.shape
        b, c, t, f = est.shape
        out = torch.empty(b * c, device=eng.device)
        with torch.cuda.device(eng.device):
            eng._ck(eng.lib.vf_lsd(eng.ctx, _ptr(est), _ptr(target), b * c, t, f, _ptr(out), _stream()))
        return out.view(b, c, 1, 1)

    def sispec(self, est: torch.Tensor, target: torch.Tensor, est_map: int = 0, target_map: int = 0) -> torch.Tensor:
        """metrics.py:89-95: scalar = sum_b sp_loss[b] / B.  est_map / target_map fuse to_log (1) / from_log (2) of the
        operands (handler() passes to_log(target_mel) and from_log(out_model['mel']), eval_gsr_voicefixer.py:60-62).
        energy_unify's pow_norm sums per (batch, channel) and pow_p_norm per batch item (utils.py:81-101): identical for
        the single-channel tensors of this path, which is what is built."""
        eng = self._eng()
        est, target = _check_in(est, eng.device, "est"), _check_in(target, eng.device, "target")
        assert est.dim() == 4 and est.shape == target.shape and est.shape[1] == 1, "sispec: [B, 1, T, F] tensors"
        b = est.shape[0]
        n = est[0].numel()
        out = torch.empty(b, device=eng.device)
        with torch.cuda.device(eng.device):
            eng._ck(eng.lib.vf_sispec(eng.ctx, _ptr(est), _ptr(target), b, n, int(est_map), int(target_map), _ptr(out), _stream()))
        return torch.sum(out) / b
