"""GPU diagnostic for the K = 256 decoder GEMMs: where do 25088 x N x 256 launches spend their time?

For the decoder shapes (M = 25088; N = 256 / 768 / 1024 / 2128) the script sweeps K = 64 .. 1024 and fits
time = fixed + per_k_block * (K / 64): `fixed` is launch + pipeline fill + the epilogue (TMEM drain, bf16 pack, TMA store of
M x N), `per_k_block` the TMA + MMA mainloop.  It prints, per N and GEMM variant, the fit next to the two floors
  * tensor : 2 M N K / sustained bf16 peak            * hbm : (M K + N K + M N) * 2 B / copy peak
and the cuBLAS time of the same call, so that the next optimisation (B-resident tiles, grouped launches, a leaner
epilogue) is chosen from data.  L2 is flushed between timed launches.

    python scripts/gpu_diag_gemm_smallk.py            # ~15 s on a B200"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimae_b200 import _lib as L  # noqa: E402
from multimae_b200 import kernels as KN  # noqa: E402


def time_us(fn, flush, iters=10):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3


def fit(ks, ts):
    """least squares t = a + b * (k / 64)"""
    xs = [k / 64.0 for k in ks]
    n = len(xs)
    mx, mt = sum(xs) / n, sum(ts) / n
    b = sum((x - mx) * (t - mt) for x, t in zip(xs, ts)) / sum((x - mx) ** 2 for x in xs)
    return mt - b * mx, b


def main():
    peaks = {"bf16_tflops_sustained": 1436.1, "hbm_gbs": 6575.4}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peaks.update(json.load(open(path)))
    tf, gbs = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = L.lib()
    M = 25088
    ks = [64, 128, 256, 512, 1024]
    for N in (256, 768, 1024, 2128):
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        bias = torch.randn(N, device=dev)
        rows = {}
        for variant in (-1, 0, 1, 3, 2, 6, 5, 4):        # -1: heuristic; 0: one tile per CTA; 1/3/2: persistent BN 128/192/256; 6/5/4: CTA pairs
            lib.mmae_gemm_set_variant(variant)
            ts = []
            try:
                for K in ks:
                    A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
                    B = (torch.randn(N, K, device=dev) * 0.5).to(torch.bfloat16)
                    ts.append(time_us(lambda: KN.gemm(A, B, bias=bias, out_bf16=out), flush))
            except L.MmaeError as e:                       # a variant that does not support the shape
                rows[variant] = "unsupported (%s)" % str(e)[:60]
                continue
            a, b = fit(ks, ts)
            rows[variant] = "K=256: %6.1f us | fit fixed %5.1f us + %5.2f us per 64-wide k-block | %s" % (
                ts[2], a, b, " ".join("%d:%.1f" % (k, t) for k, t in zip(ks, ts)))
        lib.mmae_gemm_set_variant(-1)
        K = 256
        A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        B = (torch.randn(N, K, device=dev) * 0.5).to(torch.bfloat16)
        cublas = time_us(lambda: torch.nn.functional.linear(A, B, bias.to(torch.bfloat16)), flush)
        t_tensor = 2.0 * M * N * K / (tf * 1e12) * 1e6
        t_hbm = (M * K + N * K + M * N) * 2.0 / (gbs * 1e9) * 1e6
        print("== M=%d N=%d (K=256 floors: tensor %.1f us, hbm %.1f us; cuBLAS+bias %.1f us)" % (M, N, t_tensor, t_hbm, cublas))
        for variant, line in rows.items():
            print("   variant %2d  %s" % (variant, line))


if __name__ == "__main__":
    main()
