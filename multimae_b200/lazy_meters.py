"""Host-sync removal for the UNCHANGED train_one_epoch (run_pretraining_multimae.py:525-558).

The script reads ~10 scalars per step with `.item()` (the summed loss, every task loss twice, the loss scale, the gradient
norm) and then calls `torch.cuda.synchronize()`: the host cannot enqueue step i+1 before step i has drained, so every step
starts on an empty GPU queue.  Opt-in (MMAE_LAZY_METERS=1, overlay only) the values are read ONE STEP LATE instead:

  * the criteria and the scaler return `DeferredScalar` tensors (a torch.Tensor subclass that behaves like the plain 0-dim
    tensor in every operation, including backward);
  * `DeferredScalar.item()` enqueues an asynchronous copy of the value into a pinned host slot and returns what the SAME call
    site read in the previous step (first step: the real value, with a sync);
  * the script's `torch.cuda.synchronize()` marks the step boundary and waits only for the copies of the previous step.

What the logger prints (and the `math.isfinite(loss_value)` divergence check, run_pretraining_multimae.py:529-531) therefore
lags the computation by exactly one step; nothing else changes.  Off by default: the reference semantics are the default."""
import os
import sys

import torch


class _Meters:
    def __init__(self):
        self.enabled = os.environ.get("MMAE_LAZY_METERS", "0") == "1"
        self.reset()

    def reset(self):
        self.step = 0
        self.seen = {}           # call site -> how often it has read inside the current step window
        self.slots = {}          # (call site, occurrence) -> [pinned buffer of 2 floats, [event, event], [valid, valid]]

    def _slot(self, key):
        slot = self.slots.get(key)
        if slot is None:
            slot = self.slots[key] = [torch.zeros(2, dtype=torch.float32).pin_memory(),
                                      [torch.cuda.Event(), torch.cuda.Event()], [False, False]]
        return slot

    def read(self, value, site):
        """value: 0-dim CUDA tensor; site: the reading call site (source file + line, or a name).  Returns a float: what the
        same call site (and occurrence, for a line that reads several tensors per step) read in the previous step."""
        occ = self.seen.get(site, 0)
        self.seen[site] = occ + 1
        cur = self.step & 1
        buf, events, valid = self._slot((site, occ))
        buf[cur:cur + 1].copy_(value.detach().reshape(1).float(), non_blocking=True)
        events[cur].record()
        valid[cur] = True
        prev = cur ^ 1
        if not valid[prev]:                     # first step (or a new call site): the real value
            events[cur].synchronize()
            return float(buf[cur])
        events[prev].synchronize()              # copies of the previous step: long complete
        return float(buf[prev])

    def end_step(self):
        self.step += 1
        self.seen = {}


METERS = _Meters()


class DeferredScalar(torch.Tensor):
    """A tensor whose `.item()` answers one step late when lazy meters are on (see module docstring)."""

    @staticmethod
    def wrap(t):
        if not METERS.enabled or not isinstance(t, torch.Tensor) or not t.is_cuda or isinstance(t, DeferredScalar):
            return t
        return t.as_subclass(DeferredScalar)

    def item(self):
        if METERS.enabled and self.is_cuda and self.numel() == 1:
            f = sys._getframe(1)                                   # the line of the script / logger that reads the value
            return METERS.read(self.as_subclass(torch.Tensor), (f.f_code.co_filename, f.f_lineno))
        return self.as_subclass(torch.Tensor).item()


def install(enabled=None):
    """Switch lazy meters on/off (default: MMAE_LAZY_METERS) and, when on, turn the script's per-step
    `torch.cuda.synchronize()` into the step boundary of the deferred reads."""
    if enabled is not None:
        METERS.enabled = bool(enabled)
    METERS.reset()
    if not METERS.enabled or getattr(torch.cuda.synchronize, "_mmae_lazy", False):
        return METERS.enabled
    original = torch.cuda.synchronize

    def synchronize(*args, **kwargs):
        if METERS.enabled:
            METERS.end_step()                   # the device is NOT drained: step i+1 is enqueued behind step i
            return None
        return original(*args, **kwargs)

    synchronize._mmae_lazy = True
    synchronize._mmae_original = original
    torch.cuda.synchronize = synchronize
    return True


def uninstall():
    METERS.enabled = False
    sync = torch.cuda.synchronize
    if getattr(sync, "_mmae_lazy", False):
        torch.cuda.synchronize = sync._mmae_original
