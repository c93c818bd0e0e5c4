"""GPU: validate and time the EXPERIMENTAL variant 2 of mmae_standardize_depth (multimae_b200/csrc/depth_standardize_v2.cu,
written after round 1's GPU budget was spent) against the oracle and against the validated variant 1.

    python scripts/gpu_check_depth_standardize_v2.py

Exit code 0 only if every case of tests/test_cuda_kernels._depth_cases (plus cluster-split sizes) matches the sort-based
oracle (run_pretraining_multimae.py:487-492) within the tolerances of the variant-1 test.  Flip the default in
depth_standardize.cu (depth_std_variant) only after this passes on a B200 and the timing below beats variant 1."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from multimae_b200 import _lib as L  # noqa: E402
from multimae_b200 import functional as Fn  # noqa: E402
from oracle import multimae_oracle as O  # noqa: E402  (checker only)
from test_cuda_kernels import _depth_cases  # noqa: E402


def cases():
    yield from _depth_cases()
    g = torch.Generator().manual_seed(5)
    yield "2-CTA cluster (300x300)", torch.randn(3, 1, 300, 300, generator=g).abs() + 0.1
    yield "8-CTA cluster (640x640)", torch.randn(2, 1, 640, 640, generator=g) * 2
    yield "odd length over a cluster", torch.randn(2, 1, 333, 501, generator=g)
    t = torch.round(torch.randn(2, 1, 448, 448, generator=g) * 3) / 3          # ties across both cuts, cluster of 4
    yield "448 with ties", t


def check(variant):
    L.check(L.lib().mmae_standardize_depth_set_variant(variant))
    bad = 0
    for name, x in cases():
        ref = O.standardize_depth(x)
        flat = x.reshape(x.shape[0], -1)
        n = flat.shape[1]
        trunc = torch.sort(flat, dim=1)[0][:, int(0.1 * n):int(0.9 * n)]
        out, stats = Fn.standardize_depth(x.cuda(), return_stats=True)
        torch.cuda.synchronize()
        scale = max(1.0, float(ref.abs().max()))
        err = float((out.cpu() - ref).abs().max()) / scale
        e_mean = float((stats[:, 0].cpu() - trunc.mean(1)).abs().max())
        e_var = float(((stats[:, 1].cpu() - trunc.var(1)).abs() / (trunc.var(1).abs() + 1e-7)).max())
        ok = err < 2e-5 and e_mean < 1e-5 * max(1.0, float(trunc.mean(1).abs().max())) and e_var < 1e-4
        bad += 0 if ok else 1
        print("variant %d  %-34s max|out-ref|/scale %.2e  |mean err| %.2e  rel var err %.2e  %s"
              % (variant, name, err, e_mean, e_var, "ok" if ok else "FAIL"))
    return bad


def time_us(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / iters * 1e3


def main():
    fails = check(1) + check(2)
    for B, S in ((128, 224), (32, 448)):
        x = torch.randn(B, 1, S, S, device="cuda").abs() * 3 + 0.5
        out = torch.empty_like(x)
        for variant in (1, 2):
            L.check(L.lib().mmae_standardize_depth_set_variant(variant))
            us = time_us(lambda: Fn.standardize_depth(x, out=out))
            print("time variant %d  %dx%dx%d: %.1f us = %.0f GB/s of the algorithmic 8 B/pixel"
                  % (variant, B, S, S, us, 8.0 * x.numel() / (us * 1e-6) / 1e9))
    L.check(L.lib().mmae_standardize_depth_set_variant(1))
    print("FAILS %d" % fails)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
