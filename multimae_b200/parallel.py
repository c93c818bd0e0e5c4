"""Data-parallel gradient exchange for the flat gradient arena (the only collective on the path: SURVEY.md §2.3 C1).

One process per GPU; parameters replicated; the arena is cut into contiguous buckets ordered by backward completion
(decoders -> encoder 11..0 -> input adapters).  As soon as the backward Functions report a bucket complete, its
`all_reduce(SUM)` is enqueued asynchronously (NCCL runs it on its own stream, overlapping the remaining backward
kernels); `finish()` joins them before the fused unscale/norm kernel, which also applies the 1/world_size averaging.
Replaces DistributedDataParallel's bucket copy-in/copy-out + unused-parameter bookkeeping
(run_pretraining_multimae.py:380-383) — the kernels already write gradients in place in the buckets."""
import os

import torch
import torch.distributed as dist

from . import functional as Fn


class FlatGradReducer:
    def __init__(self, arena, param_names_in_registration_order, process_group=None, bucket_bytes=48 << 20):
        self.arena = arena
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # buckets: walk parameters in reverse registration order (= backward completion order)
        self.buckets = []       # [lo, hi, set(names)]
        cur_names, hi, lo = set(), None, None
        for name in reversed(list(param_names_in_registration_order)):
            o, n, _ = arena.offsets[name]
            end = o + (n + 7) // 8 * 8
            if hi is None:
                hi = end
            lo = o
            cur_names.add(name)
            if (hi - lo) * 4 >= bucket_bytes:
                self.buckets.append([lo, hi, cur_names])
                cur_names, hi, lo = set(), None, None
        if cur_names:
            self.buckets.append([lo, hi, cur_names])
        self.bucket_of = {n: i for i, b in enumerate(self.buckets) for n in b[2]}
        self._pending = None
        self._works = []
        self.reset()

    def reset(self):
        self._pending = [len(b[2]) for b in self.buckets]
        self._works = []
        self._events = [[] for _ in self.buckets]      # completion events of producers that ran on other streams

    def on_grads_ready(self, names):
        """Called from backward (after the producing kernels were enqueued on the current stream)."""
        cuda = self.arena.flat.is_cuda
        touched = {self.bucket_of[n] for n in names if n in self.bucket_of}
        if cuda and self.world > 1:
            # the task decoders run their backward on their own streams: the reduction of a bucket has to follow every
            # producer stream, not only the one that happens to complete the bucket
            ev = torch.cuda.current_stream().record_event()
            for i in touched:
                self._events[i].append(ev)
        for n in names:
            i = self.bucket_of.get(n)
            if i is None:
                continue
            self._pending[i] -= 1
            if self._pending[i] == 0 and self.world > 1:
                lo, hi, _ = self.buckets[i]
                if cuda:
                    cur = torch.cuda.current_stream()
                    for e in self._events[i]:
                        cur.wait_event(e)
                self._works.append(dist.all_reduce(self.arena.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                                   async_op=True))

    def finish(self):
        """Join outstanding reductions; returns the factor that turns the summed gradients into the average."""
        if self.world > 1:
            missing = [i for i, p in enumerate(self._pending) if p > 0]
            for i in missing:                      # parameters that received no gradient this step (still exchanged)
                lo, hi, _ = self.buckets[i]
                if self.arena.flat.is_cuda:
                    for e in self._events[i]:
                        torch.cuda.current_stream().wait_event(e)
                self._works.append(dist.all_reduce(self.arena.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                                   async_op=True))
            for w in self._works:
                w.wait()
        self.reset()
        return 1.0 / self.world


def attach_data_parallel(model, scaler=None, process_group=None, bucket_bytes=None):
    """Wire a MultiMAE model for data-parallel training with in-place bucketed all-reduce; returns the reducer.
    `bucket_bytes`: default 48 MB (MMAE_BUCKET_MB overrides it for sweeps)."""
    if bucket_bytes is None:
        bucket_bytes = int(float(os.environ.get("MMAE_BUCKET_MB", "48")) * (1 << 20))
    arena = model.own_gradients(True)
    names = list(arena.offsets)      # the arena's own order (registration order, proj_context tensors grouped): contiguous buckets
    reducer = FlatGradReducer(arena, names, process_group, bucket_bytes)
    arena.reducer = reducer            # found again by a scaler that only ever sees model.parameters() (overlay launcher)
    model.set_grad_callback(reducer.on_grads_ready)
    if scaler is not None:
        scaler.attach_arena(arena)
        scaler.attach_reducer(reducer)
    return reducer


def broadcast_parameters(model, src=0, process_group=None):
    """Replicate rank-`src` parameters (DDP's constructor broadcast, run_pretraining_multimae.py:381)."""
    if dist.is_initialized() and dist.get_world_size(process_group) > 1:
        with torch.no_grad():
            for p in model.parameters():
                dist.broadcast(p.data, src=src, group=process_group)
    Fn.invalidate_mirrors()                                   # `.data` writes are invisible to the version counters
