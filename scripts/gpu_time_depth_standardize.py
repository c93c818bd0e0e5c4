"""GPU: time mmae_standardize_depth against the reference's torch expression (run_pretraining_multimae.py:487-492) on the
same device, at the cfg-2 shape (128 x 224 x 224) and the cfg-5 shape (32 x 448 x 448).  Prints us per call and the
achieved fraction of the copy bandwidth for the algorithmic 8 B/pixel (one read + one write of the map)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimae_b200 import functional as Fn  # noqa: E402


def torch_expr(depth):
    flat = depth.reshape(depth.shape[0], -1)
    trunc = torch.sort(flat, dim=1)[0]
    trunc = trunc[:, int(0.1 * trunc.shape[1]): int(0.9 * trunc.shape[1])]
    return (depth - trunc.mean(dim=1)[:, None, None, None]) / torch.sqrt(trunc.var(dim=1)[:, None, None, None] + 1e-6)


def time_us(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    total = 0.0
    for _ in range(iters):
        flush.zero_()                                  # > L2: the map comes from HBM every time
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / iters * 1e3


def main():
    peak = 6575.4
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak = json.load(open(path)).get("hbm_gbs", peak)
    for B, S in ((128, 224), (32, 448)):
        x = (torch.randn(B, 1, S, S, device="cuda").abs() * 3 + 0.5)
        out = torch.empty_like(x)
        ours = time_us(lambda: Fn.standardize_depth(x, out=out))
        ref = time_us(lambda: torch_expr(x), iters=5)
        err = float((out - torch_expr(x)).abs().max())
        gbs = 8.0 * x.numel() / (ours * 1e-6) / 1e9
        print("standardize_depth %dx%dx%d: kernel %.1f us (%.0f GB/s of the algorithmic 8 B/pixel = %.2f of the %.0f GB/s copy "
              "peak); torch sort expression %.1f us; max |diff| %.2e" % (B, S, S, ours, gbs, gbs / peak, peak, ref, err))


if __name__ == "__main__":
    main()
