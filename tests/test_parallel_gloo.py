"""CPU, world_size 2 over gloo: the bucketed in-place gradient exchange (multimae_b200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimae_b200.functional import GradArena
from multimae_b200.parallel import FlatGradReducer


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    named = [("a.w", torch.zeros(37, 5)), ("a.b", torch.zeros(5)), ("b.w", torch.zeros(64, 64)), ("b.b", torch.zeros(3)),
             ("c.w", torch.zeros(129))]
    arena = GradArena(named, torch.device("cpu"))
    red = FlatGradReducer(arena, [n for n, _ in named], bucket_bytes=4096)
    assert len(red.buckets) >= 2 and sorted(n for b in red.buckets for n in b[2]) == sorted(n for n, _ in named)
    # buckets tile the arena in reverse registration order without overlap
    spans = sorted((b[0], b[1]) for b in red.buckets)
    assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    for step in range(2):
        arena.zero_()
        for i, (n, _) in enumerate(named):
            arena.view(n).fill_(float(rank + 1) * (i + 1) + step)
        for n in ["c.w", "b.b", "b.w"]:              # backward order; 'a.*' never reported -> flushed by finish()
            red.on_grads_ready([n])
        scale = red.finish()
        assert scale == 1.0 / world
        for i, (n, _) in enumerate(named):
            # the arena holds the SUM; `scale` (1/world) is folded into the fused unscale/norm pass by the scaler
            expect = sum(float(r + 1) * (i + 1) + step for r in range(world))
            assert torch.allclose(arena.view(n), torch.full_like(arena.view(n), expect)), (n, step)
    out.put((rank, "ok"))
    dist.destroy_process_group()


def test_flat_grad_reducer_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [(0, "ok"), (1, "ok")]
