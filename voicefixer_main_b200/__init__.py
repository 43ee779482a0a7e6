"""B200-native VoiceFixer inference hot path (stages A/B/C of eval_gsr_voicefixer.py:handler)."""
from .arch import VocoderConfig  # noqa: F401
from .model import (Engine, FDomainHelper, HParams, MelScale, VoiceFixer, Vocoder,  # noqa: F401
                    default_hparams, get_hparams_from_file)
