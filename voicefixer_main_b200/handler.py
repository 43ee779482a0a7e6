"""Mirror of eval_gsr_voicefixer.py (pre :19-25, refresh_model :31-35, handler :37-77) over the B200 engine.

Same signature and contract as the reference handler that evaluation_proc/eval.py:128-132 calls:
    handler(input, output, target, ckpt, device, needrefresh=False, meta={}) -> dict of metrics
It writes a 16-bit wav at `output`.  Differences, all outside the hot path: audio I/O uses the stdlib `wave`
module (librosa / soundfile are not in this image, so inputs must already be 44.1 kHz PCM16 wav), and the mel
metrics that need `target` (evaluation_proc/metrics.py, CPU-side, out of scope) are not computed.
The segment loop, from_log, peak normalisation, trim_center, concat and the int16 conversion of
tools/file/wav.py:22-24 are reproduced exactly; the per-segment stages run as one fused launch chain.
"""
import wave

import numpy as np
import torch

from .model import VoiceFixer, default_hparams

model = None
hp = None
SEG_LENGTH = 44100 * 60          # eval_gsr_voicefixer.py:47


def load_wav(path, sample_rate=44100):
    """tools/utils.py:46-48 for PCM16 input at the target rate (no resampling available here)."""
    with wave.open(path, "rb") as w:
        if w.getframerate() != sample_rate or w.getsampwidth() != 2:
            raise ValueError(f"{path}: need {sample_rate} Hz 16-bit PCM (got {w.getframerate()} Hz, {8 * w.getsampwidth()} bit)")
        data = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).reshape(-1, w.getnchannels())
    return (data.astype(np.float32) / 32768.0).mean(axis=1).astype(np.float32)


def save_wave(frames: np.ndarray, fname, sample_rate=44100):
    """tools/file/wav.py:10-27: scale by 2^15 when max <= 1 and truncate toward zero to int16."""
    frames = np.array(frames, dtype=np.float32, copy=True).reshape(-1)
    if np.max(frames) <= 1:
        frames *= 2 ** 15
    pcm = frames.astype(np.short)
    with wave.open(fname, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm.tobytes())


def save_pcm16(pcm: np.ndarray, fname, sample_rate=44100):
    """Write already-converted int16 samples (VoiceFixer.restore_pcm16 / Engine.to_pcm16)."""
    with wave.open(fname, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1).tobytes())


def refresh_model(ckpt):
    global model
    model = VoiceFixer(hp if hp is not None else default_hparams(), channels=2, type_target="vocals").load_from_checkpoint(ckpt)
    model.eval()


def restore_array(mdl: VoiceFixer, wav_10k: np.ndarray, device, unify_energy: bool = False) -> torch.Tensor:
    """The segment loop of handler() for one in-memory file: returns [1, N] on `device`."""
    res = []
    break_point = SEG_LENGTH
    n = wav_10k.shape[0]
    while break_point < n + SEG_LENGTH:
        segment = wav_10k[break_point - SEG_LENGTH:break_point]
        seg = torch.from_numpy(np.ascontiguousarray(segment))[None, :].to(device)
        res.append(mdl.restore(seg, unify_energy=unify_energy))
        break_point += SEG_LENGTH
    return torch.cat(res, -1)


def handler(input, output, target, ckpt, device, needrefresh=False, meta={}):
    if needrefresh:
        refresh_model(ckpt)
    global model
    model = model.to(device)
    metrics = {}
    wav_10k = load_wav(input, sample_rate=44100)
    out = restore_array(model, wav_10k, model.device, unify_energy=bool(meta.get("unify_energy", False)))
    # save_wave's float -> int16 conversion runs on the GPU (its `max <= 1` condition always holds after the
    # per-segment peak normalisation), so only 2 bytes per sample cross PCIe
    save_pcm16(model._engine().to_pcm16(out[0]).cpu().numpy(), fname=output, sample_rate=44100)
    return metrics
