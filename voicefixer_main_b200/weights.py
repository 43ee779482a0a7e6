"""Deterministic synthetic weights in the reference's state-dict naming.

No trained checkpoint or vocoder weights can be fetched here (no network), so
benchmarks and parity tests use seeded weights: xavier-uniform convs as
modules.py:276-282 initialises them, and *non-trivial* BatchNorm statistics
 so the eval-mode BN fold is actually exercised.  The same
generator feeds the reference model (oracle side, via load_state_dict) and the
CUDA engine, so both see bit-identical fp32 parameters.
"""
import math
from typing import Dict

import torch

from .arch import SSR_PREFIX, UNET_PREFIX, VocoderConfig, unet_keys, vocoder_keys


def _xavier(shape, gen, transposed=False):
    # nn.init.xavier_uniform_: fan_in = size(1)*rf, fan_out = size(0)*rf
    rf = 1
    for s in shape[2:]:
        rf *= s
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound


def make_unet_state(seed: int = 1234) -> Dict[str, torch.Tensor]:
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in unet_keys():
        key = UNET_PREFIX + name
        leaf = name.rsplit(".", 1)[1]
        parent = name.rsplit(".", 2)[-2]
        if leaf == "num_batches_tracked":
            sd[key] = torch.zeros((), dtype=torch.long)
        elif parent.startswith("bn"):
            if leaf == "weight":
                sd[key] = torch.rand(shape, generator=gen) * 0.5 + 0.75
            elif leaf == "bias":
                sd[key] = torch.randn(shape, generator=gen) * 0.1
            elif leaf == "running_mean":
                sd[key] = torch.randn(shape, generator=gen) * 0.1
            else:
                sd[key] = torch.rand(shape, generator=gen) * 0.5 + 0.75
        elif leaf == "weight":
            sd[key] = _xavier(shape, gen)
        else:  # conv biases (shortcut / head); small but non-zero so the bias path is tested
            sd[key] = torch.randn(shape, generator=gen) * 0.05
    return sd


def make_vocoder_state(cfg: VocoderConfig = None, seed: int = 4321) -> Dict[str, torch.Tensor]:
    """PyTorch default conv init (U(+-1/sqrt(fan_in))) keeps the 70-layer
    residual generator numerically tame with random weights."""
    cfg = cfg or VocoderConfig()
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in vocoder_keys(cfg):
        if name.endswith("weight"):
            if name.startswith("up."):
                fan_in = shape[0] * 2         # two taps of the transposed conv reach each output sample
            else:
                fan_in = shape[1] * shape[2]
            bound = 1.0 / math.sqrt(fan_in)
            sd["vocoder." + name] = (torch.rand(shape, generator=gen) * 2 - 1) * bound
        else:
            sd["vocoder." + name] = (torch.rand(shape, generator=gen) * 2 - 1) * 0.05
    return sd


def make_state(seed: int = 1234, cfg: VocoderConfig = None) -> Dict[str, torch.Tensor]:
    sd = make_unet_state(seed)
    sd.update(make_vocoder_state(cfg, seed + 1))
    return sd


def make_ssr_state(seed: int = 1234) -> Dict[str, torch.Tensor]:
    """unet_v2 (models/components/unet_v2.py) has the mel UNet's parameter shapes: the same seeded tensors under the
    SSR_UNet / GSR_UNet prefix (models/ssr_unet.py:49)."""
    return {k.replace(UNET_PREFIX, SSR_PREFIX): v for k, v in make_unet_state(seed).items()}
