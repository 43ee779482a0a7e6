"""world_size-2 gloo tests of the multi-GPU plumbing (weight broadcast + clip sharding) on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voicefixer_main_b200 import dist as vdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = vdist.init_from_env(backend="gloo")
    layout = [("a.weight", (4, 3), 12), ("b.bias", (5,), 5)]
    state = {"a.weight": torch.arange(12.).view(4, 3), "b.bias": torch.full((5,), 7.0)} if r == 0 else None
    got = vdist.broadcast_state(state, layout, torch.device("cpu"))
    ok = torch.equal(got["a.weight"], torch.arange(12.).view(4, 3)) and torch.equal(got["b.bias"], torch.full((5,), 7.0))
    lo, hi = vdist.shard_range(7, r, w)
    mx = vdist.max_over_ranks(float(r + 1), torch.device("cpu"))
    q.put((r, ok, lo, hi, mx))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_sharding_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == (0, True, 0, 4, 2.0) and res[1] == (1, True, 4, 7, 2.0)


def test_shard_range_covers_everything():
    for total in (1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [vdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_layout_from_arch_matches_synthetic_state():
    from voicefixer_main_b200.weights import make_state
    sd = make_state(1)
    layout = vdist.layout_from_arch()
    assert all(k in sd and tuple(sd[k].shape) == s and sd[k].numel() == n for k, s, n in layout)
    assert sum(n for _, _, n in layout) > 90_000_000     # UNet 65 M + vocoder ~34 M parameters
