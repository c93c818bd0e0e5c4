"""GPU bring-up check for the tcgen05 GEMM and element-wise kernels (run under gpurun).

Prints one line per case; exits non-zero if any case fails.  Every case is bounded; run under `timeout`.
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from multimae_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
fails = 0


def relerr(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def report(name, err, tol):
    global fails
    ok = err < tol
    fails += 0 if ok else 1
    print("%-58s relerr=%.3e  %s" % (name, err, "ok" if ok else "FAIL"), flush=True)


def rand_bf16(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).to(torch.bfloat16)


def gemm_case(M, N, K, a_mn, b_mn, split_k=1):
    A = rand_bf16(M, K)
    B = rand_bf16(N, K)
    ref = A.float() @ B.float().t()
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    out = torch.zeros(M, N, device=dev, dtype=torch.float32)
    K.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, out_f32=out, split_k=split_k)
    torch.cuda.synchronize()
    report("gemm M=%d N=%d K=%d a_mn=%d b_mn=%d split=%d" % (M, N, K, a_mn, b_mn, split_k), relerr(out, ref), 1e-5)


# ---- correctness: the four operand-major combinations, small and ragged shapes
for (M, N, K_) in [(128, 128, 64), (128, 128, 256), (256, 384, 768), (200, 136, 200), (396, 2128, 256)]:
    for a_mn in (False, True):
        for b_mn in (False, True):
            if a_mn and M % 8:
                continue
            try:
                gemm_case(M, N, K_, a_mn, b_mn)
            except Exception as e:  # noqa: BLE001
                fails += 1
                print("gemm M=%d N=%d K=%d a_mn=%d b_mn=%d EXC %s" % (M, N, K_, a_mn, b_mn, e), flush=True)

# split-K (wgrad shape): dW[768,768] = dY^T X with contraction 12672
for split in (1, 4):
    gemm_case(768, 768, 12672, True, True, split_k=split)
gemm_case(768, 3072, 1280, True, True, split_k=3)

# ---- fused epilogues
M, N, K_ = 384, 512, 256
A = rand_bf16(M, K_)
B = rand_bf16(N, K_)
bias = torch.randn(N, device=dev)
resid = torch.randn(M, N, device=dev)
acc = A.float() @ B.float().t()

out = torch.empty(M, N, device=dev)
K.gemm(A, B, bias=bias, out_f32=out)
report("epilogue bias", relerr(out, acc + bias), 1e-5)

outb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
K.gemm(A, B, bias=bias, act=1, preact=pre, out_bf16=outb)
z = acc + bias
report("epilogue bias+gelu -> bf16", relerr(outb, torch.nn.functional.gelu(z)), 4e-3)
report("epilogue preact bf16", relerr(pre, z), 4e-3)

out = torch.empty(M, N, device=dev)
K.gemm(A, B, bias=bias, residual=resid, out_f32=out)
report("epilogue bias+residual", relerr(out, acc + bias + resid), 1e-5)

zz = (torch.randn(M, N, device=dev)).to(torch.bfloat16)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
K.gemm(A, B, dgelu_z=zz, out_bf16=out)
zf = zz.float().requires_grad_(True)
torch.nn.functional.gelu(zf).sum().backward()
report("epilogue dgelu", relerr(out, acc * zf.grad), 4e-3)

out = torch.ones(M, N, device=dev)
K.gemm(A, B, out_f32=out, accumulate=True, alpha=0.5)
report("epilogue accumulate alpha", relerr(out, 1 + 0.5 * acc), 1e-5)

# ---- element-wise kernels
x = torch.randn(1000, 776, device=dev)
report("cast_bf16", relerr(K.cast_bf16(x), x.to(torch.bfloat16)), 1e-7)
dst = torch.empty(1000, 776, device=dev, dtype=torch.bfloat16)
cs = torch.zeros(776, device=dev)
K.cast_colsum(x, dst, cs)
report("cast_colsum cast", relerr(dst, x.to(torch.bfloat16)), 1e-7)
report("cast_colsum sum", relerr(cs, x.sum(0)), 1e-5)
cs2 = torch.zeros(776, device=dev)
K.colsum_bf16(dst, cs2)
report("colsum_bf16", relerr(cs2, dst.float().sum(0)), 1e-5)
t = K.transpose_bf16(dst)
report("transpose_bf16", relerr(t, dst.t()), 1e-7)

# ---- quick timing vs cuBLAS on the encoder shapes (device events, inputs > L2 not enforced here: indicative)
def time_it(fn, iters=20):
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


for (M, N, K_, a_mn, b_mn, split) in [(12672, 2304, 768, 0, 0, 1), (12672, 768, 768, 0, 0, 1),
                                      (12672, 3072, 768, 0, 0, 1), (12672, 768, 3072, 0, 0, 1),
                                      (12672, 768, 3072, 0, 1, 1), (3072, 768, 12672, 1, 1, 2),
                                      (768, 768, 12672, 1, 1, 8), (25088, 1024, 256, 0, 0, 1)]:
    A = rand_bf16(M, K_)
    B = rand_bf16(N, K_)
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    if split > 1:
        out = torch.zeros(M, N, device=dev)
        ms = time_it(lambda: K.gemm(Ain, Bin, a_mn=bool(a_mn), b_mn=bool(b_mn), out_f32=out, split_k=split))
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ms = time_it(lambda: K.gemm(Ain, Bin, a_mn=bool(a_mn), b_mn=bool(b_mn), out_bf16=out))
    ms_ref = time_it(lambda: torch.matmul(A, B.t()))
    fl = 2.0 * M * N * K_
    print("time M=%d N=%d K=%d a_mn=%d b_mn=%d split=%d : ours %.3f ms (%.0f TF/s)  cublas %.3f ms (%.0f TF/s)" %
          (M, N, K_, a_mn, b_mn, split, ms, fl / ms / 1e9, ms_ref, fl / ms_ref / 1e9), flush=True)

print("FAILS", fails)
sys.exit(1 if fails else 0)
