"""Host-side pieces of the handler mirror (no GPU): wav I/O semantics of tools/file/wav.py:22-24."""
import numpy as np

from voicefixer_main_b200 import handler as H


def test_save_wave_truncates_like_reference(tmp_path):
    x = np.array([0.5, -0.5, 0.99999, -0.00002, 0.0], dtype=np.float32)
    p = str(tmp_path / "o.wav")
    H.save_wave(x, p)
    back = H.load_wav(p)
    assert (np.round(back * 32768).astype(int) == np.array([16384, -16384, 32767, 0, 0])).all()


def test_segment_loop_bounds():
    # same loop arithmetic as eval_gsr_voicefixer.py:47-50: ceil(n / 60 s) segments, last one ragged
    seg = H.SEG_LENGTH
    for n in (1, seg - 1, seg, seg + 1, 3 * seg + 5):
        cnt, bp = 0, seg
        while bp < n + seg:
            cnt += 1
            bp += seg
        assert cnt == -(-n // seg)
