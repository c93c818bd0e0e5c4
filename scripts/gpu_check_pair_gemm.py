"""GPU: CTA-pair GEMM kernels (variants 4/5/6) against fp32 matmul, then timing against the single-CTA persistent kernels.
Run under a short `timeout`: a protocol error between the two CTAs of a pair shows up as a hang."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from helpers import rel_l2  # noqa: E402
from multimae_b200 import _lib as L  # noqa: E402
from multimae_b200 import kernels as KN  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
lib = L.lib()


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)


fails = 0
for variant in (4, 5, 6):
    lib.mmae_gemm_set_variant(variant)
    for (M, N, K) in [(256, 256, 64), (256, 256, 256), (512, 768, 512), (200, 136, 200), (396, 2128, 256), (1000, 776, 512),
                      (2560, 2304, 768), (12672, 768, 768)]:
        A, B = bf(M, K), bf(N, K)
        ref = A.float() @ B.float().t()
        for a_mn in (False, True):
            for b_mn in (False, True):
                if a_mn and M % 8:
                    continue
                out = torch.zeros(M, N, device=dev)
                KN.gemm(A.t().contiguous() if a_mn else A, B.t().contiguous() if b_mn else B, a_mn=a_mn, b_mn=b_mn, out_f32=out)
                torch.cuda.synchronize()
                err = rel_l2(out, ref)
                ok = err < 3e-5
                fails += 0 if ok else 1
                print("variant %d  %5dx%5dx%5d a_mn=%d b_mn=%d  relerr %.2e %s" % (variant, M, N, K, a_mn, b_mn, err, "ok" if ok else "FAIL"),
                      flush=True)
        bias = torch.randn(N, device=dev)
        ob = torch.full((M + 3, N + 16), 7.0, device=dev, dtype=torch.bfloat16)
        KN.gemm(A, B, bias=bias, out_bf16=ob[:M, :N])
        e1 = rel_l2(ob[:M, :N], ref + bias)
        clean = bool((ob[M:] == 7).all()) and bool((ob[:, N:] == 7).all())
        acc = torch.ones(M, N, device=dev)
        KN.gemm(A, B, out_f32=acc, accumulate=True, split_k=3)
        e2 = rel_l2(acc, 1 + ref)
        ok = e1 < 4e-3 and e2 < 3e-5 and clean
        fails += 0 if ok else 1
        print("variant %d  %5dx%5dx%5d bias->bf16 %.2e  split-3 accumulate %.2e  margins clean %s  %s" %
              (variant, M, N, K, e1, e2, clean, "ok" if ok else "FAIL"), flush=True)
lib.mmae_gemm_set_variant(-1)
print("PAIR GEMM FAILS", fails, flush=True)


def time_it(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (M, N, K, a_mn, b_mn, split) in [(12672, 3072, 768, 0, 0, 1), (12672, 3072, 768, 0, 1, 1), (12672, 768, 3072, 0, 0, 1),
                                     (12672, 768, 3072, 0, 1, 1), (12672, 2304, 768, 0, 0, 1), (12672, 768, 2304, 0, 1, 1),
                                     (12672, 768, 768, 0, 0, 1), (3072, 768, 12672, 1, 1, 1), (768, 3072, 12672, 1, 1, 1),
                                     (2304, 768, 12672, 1, 1, 1), (25088, 1024, 256, 0, 0, 1), (25088, 256, 1024, 0, 0, 1)]:
    A = bf(K, M) if a_mn else bf(M, K)
    B = bf(K, N) if b_mn else bf(N, K)
    bias = torch.randn(N, device=dev)
    fl = 2.0 * M * N * K
    res = []
    for variant in (-1, 2, 3, 4, 5, 6):
        lib.mmae_gemm_set_variant(variant)
        if a_mn:
            out = torch.zeros(M, N, device=dev)
            ms = time_it(lambda: KN.gemm(A, B, a_mn=True, b_mn=bool(b_mn), out_f32=out, accumulate=True, split_k=split))
        else:
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            ms = time_it(lambda: KN.gemm(A, B, b_mn=bool(b_mn), bias=bias, out_bf16=out))
        res.append("v%d %.1f us (%.0f)" % (variant, ms * 1e3, fl / ms / 1e9))
    lib.mmae_gemm_set_variant(-1)
    print("time %5dx%5dx%5d a_mn=%d b_mn=%d: %s" % (M, N, K, a_mn, b_mn, " | ".join(res)), flush=True)
