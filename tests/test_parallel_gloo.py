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


def _overlay_worker(rank, world, port, out):
    """The overlay's multi-rank path as run_pretraining_multimae.py drives it (:380-391, 536-537): the model is wrapped by
    the DistributedDataParallel stand-in, the scaler is created afterwards and only ever sees `model.parameters()`."""
    import ctypes
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimae_b200 import _lib as L
    from multimae_b200 import functional as Fn
    from multimae_b200 import overlay
    from multimae_b200.criterion import MaskedMSELoss
    from multimae_b200.native_scaler import NativeScalerWithGradNormCount
    from test_host_api import _build

    class Stub:                                     # no GPU here: validate the calls, compute nothing
        def __getattr__(self, name):
            res, argtypes = L.SIGNATURES[name]

            def fn(*args):
                assert len(args) == len(argtypes), name
                return 4096 if name.endswith("_bytes") else (L.ABI_VERSION if name == "mmae_abi_version" else 0)
            return fn
    stub = Stub()
    L.lib = lambda: stub
    L.current_stream = lambda: 0
    Fn._require_cuda = lambda t, what: None
    seen = {}

    def fake_unscale_norm(flat, inv_scale=1.0, post_scale=1.0, inv_scale_tensor=None):
        seen["post_scale"], seen["flat"] = post_scale, flat
        return torch.ones(()), torch.zeros(2)
    Fn.grad_unscale_norm = fake_unscale_norm

    torch.manual_seed(rank)                         # different initial weights per rank: the wrapper must broadcast rank 0's
    model = _build(in_domains=("rgb",)).train()
    wrapped = overlay._IdentityDDP(model, device_ids=[0], find_unused_parameters=True)
    assert wrapped.module is model and model._mmae_reducer is model.grad_arena().reducer
    w0 = model.encoder[0].attn.qkv.weight.detach().clone()
    gathered = [torch.empty_like(w0) for _ in range(world)]
    dist.all_gather(gathered, w0)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    reducer = model._mmae_reducer
    arena = model.grad_arena()
    ready = reducer.on_grads_ready

    def fill_then_report(names):                    # what the backward kernels would have written: rank + 1 everywhere
        for n in names:
            arena.view(n).fill_(float(rank + 1))
        ready(names)
    model.set_grad_callback(fill_then_report)
    optimizer = torch.optim.AdamW(model.parameters(), lr=0.0)
    scaler = NativeScalerWithGradNormCount(enabled=True)        # never told about the arena or the reducer
    x = {"rgb": torch.randn(2, 3, 64, 64)}
    preds, masks = wrapped(x, num_encoded_tokens=4)
    loss = sum(MaskedMSELoss(16, 1)(preds[k], x["rgb"], mask=masks["rgb"]) for k in preds)
    optimizer.zero_grad()
    norm = scaler(loss, optimizer, clip_grad=None, skip_grad=None, parameters=wrapped.parameters())
    assert norm is not None and seen["post_scale"] == 1.0 / world and seen["flat"] is arena.flat
    total = float(sum(r + 1 for r in range(world)))
    for n, p in model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and p.grad.data_ptr() == arena.view(n).data_ptr(), n
            assert torch.all(arena.view(n) == total), n                 # every bucket was exchanged exactly once
    assert all(c == len(b[2]) for c, b in zip(reducer._pending, reducer.buckets)) and not reducer._works
    out.put((rank, "ok"))
    dist.destroy_process_group()


def test_overlay_ddp_stand_in_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlay_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [(0, "ok"), (1, "ok")]
