"""Multi-GPU plumbing: one process per GPU, weights broadcast once, clips sharded with no data-path collective.

Utterances (and the 60 s segments of a long file) are independent end to end
(eval_gsr_voicefixer.py:49-74 keeps no cross-segment state), so the only collective on the whole path is a
single broadcast of the flattened fp32 state from rank 0 at start-up (NCCL over NVLink on GPUs, gloo in the
CPU tests).  Every rank then packs its own copy inside libb200vf and runs its contiguous slice of the batch.
"""
import os
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises the process group when
    WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def state_layout(state: Dict[str, torch.Tensor]) -> List[Tuple[str, Tuple[int, ...], int]]:
    """Deterministic (name, shape, numel) list of the floating-point tensors of a state dict."""
    return [(k, tuple(v.shape), v.numel()) for k, v in sorted(state.items()) if v.is_floating_point()]


def broadcast_state(state: Dict[str, torch.Tensor], layout, device, src: int = 0) -> Dict[str, torch.Tensor]:
    """One broadcast of the whole parameter blob.  `state` is only read on rank `src`; `layout` (from
    state_layout on any rank that can build it, e.g. from arch.py shapes) must be identical everywhere."""
    total = sum(n for _, _, n in layout)
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if not dist.is_initialized() or dist.get_rank() == src:
        off = 0
        for k, _, n in layout:
            flat[off:off + n] = state[k].reshape(-1).to(device=device, dtype=torch.float32)
            off += n
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)
    out, off = {}, 0
    for k, shape, n in layout:
        out[k] = flat[off:off + n].view(shape)
        off += n
    return out


def layout_from_arch(cfg=None):
    """The layout every rank can compute without holding the weights."""
    from .arch import UNET_PREFIX, VocoderConfig, unet_keys, vocoder_keys
    cfg = cfg or VocoderConfig()
    items = [(UNET_PREFIX + k, tuple(s)) for k, s in unet_keys() if not k.endswith("num_batches_tracked")]
    items += [("vocoder." + k, tuple(s)) for k, s in vocoder_keys(cfg)]
    out = []
    for k, s in sorted(items):
        n = 1
        for d in s:
            n *= d
        out.append((k, s, n))
    return out


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of `total` clips; the first `total % world` ranks take one extra."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
