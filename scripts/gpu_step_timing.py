"""Quick device-timed breakdown of one MultiMAE-B pre-training step on the CUDA path (run under gpurun)."""
import argparse
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from multimae_b200.criterion import MaskedCrossEntropyLoss, MaskedL1Loss, MaskedMSELoss  # noqa: E402
from multimae_b200.native_scaler import NativeScalerWithGradNormCount  # noqa: E402
from multimae_b200.optim import FlatAdamW  # noqa: E402
from test_host_api import _build  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--tiny", action="store_true")
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--variant", type=int, default=-1)
a = ap.parse_args()

dev = torch.device("cuda:0")
torch.manual_seed(0)
from multimae_b200 import _lib as _L  # noqa: E402
_L.lib().mmae_gemm_set_variant(a.variant)
if a.tiny:
    model = _build(("rgb", "depth", "semseg"), 128, 2, 2, 128, 1, 4, 64)
    size, T = 64, 12
else:
    model = _build(("rgb", "depth", "semseg"), 768, 12, 12, 256, 2, 8, 224)
    size, T = 224, 98
model = model.to(dev).train()
opt = FlatAdamW(model, lr=1e-4)
scaler = NativeScalerWithGradNormCount(enabled=False).attach_arena(model.grad_arena())
B = a.batch
x = {"rgb": torch.randn(B, 3, size, size, device=dev), "depth": torch.randn(B, 1, size, size, device=dev),
     "semseg": torch.randint(0, 133, (B, size // 4, size // 4), device=dev)}
fns = {"rgb": MaskedMSELoss(16, 1), "depth": MaskedL1Loss(16, 1), "semseg": MaskedCrossEntropyLoss(16, 4),
       "norm_rgb": MaskedMSELoss(16, 1, norm_pix=True)}


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def step(timed=False):
    marks = [ev()] if timed else None
    preds, masks = model(x, num_encoded_tokens=T, alphas=1.0)
    if timed:
        marks.append(ev())
    loss = sum(fns[k](preds[k], x["rgb" if k == "norm_rgb" else k], mask=masks["rgb" if k == "norm_rgb" else k])
               for k in preds)
    if timed:
        marks.append(ev())
    scaler(loss, opt, parameters=None)
    if timed:
        marks.append(ev())
    return loss, marks


for _ in range(a.warmup):
    loss, _m = step()
torch.cuda.synchronize()
print("warm loss", float(loss))
t0 = time.perf_counter()
e0 = ev()
for _ in range(a.steps):
    loss, _m = step()
e1 = ev()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / a.steps * 1e3
msd = e0.elapsed_time(e1) / a.steps
print("step: %.2f ms device, %.2f ms wall -> %.0f samples/s (B=%d)" % (msd, wall, B / msd * 1e3, B))
loss, marks = step(timed=True)
torch.cuda.synchronize()
names = ["forward(model)", "losses fwd", "backward+norm+adamw"]
for n, (s, e) in zip(names, zip(marks[:-1], marks[1:])):
    print("  %-22s %.2f ms" % (n, s.elapsed_time(e)))
print("final loss", float(loss), "mem GB", torch.cuda.max_memory_allocated() / 2 ** 30)
# CPU-side launch overhead: time the python side only (async)
torch.cuda.synchronize()
t0 = time.perf_counter()
loss, _m = step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue time for one step: %.2f ms" % ((t1 - t0) * 1e3))
