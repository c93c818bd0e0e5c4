"""GPU bring-up check for the tcgen05 GEMM and element-wise kernels (run under gpurun).

Prints one line per case; exits non-zero if any case fails.  Every case is bounded; run under `timeout`.
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from multimae_b200 import kernels as KN  # noqa: E402

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
    KN.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, out_f32=out, split_k=split_k)
    torch.cuda.synchronize()
    report("gemm M=%d N=%d K=%d a_mn=%d b_mn=%d split=%d" % (M, N, K, a_mn, b_mn, split_k), relerr(out, ref), 3e-5)


from multimae_b200 import _lib as LIB  # noqa: E402

# ---- correctness: the four operand-major combinations, small and ragged shapes, for every kernel variant
for variant in (0, 1, 2, 3):
    LIB.lib().mmae_gemm_set_variant(variant)
    print("== gemm variant", variant, flush=True)
    for (M, N, K_) in [(128, 128, 64), (128, 128, 256), (256, 384, 768), (200, 136, 200), (396, 2128, 256),
                       (1000, 768, 512), (2560, 2304, 768)]:
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
LIB.lib().mmae_gemm_set_variant(-1)

# ---- fused epilogues
M, N, K_ = 384, 512, 256
A = rand_bf16(M, K_)
B = rand_bf16(N, K_)
bias = torch.randn(N, device=dev)
resid = torch.randn(M, N, device=dev)
acc = A.float() @ B.float().t()

out = torch.empty(M, N, device=dev)
KN.gemm(A, B, bias=bias, out_f32=out)
report("epilogue bias", relerr(out, acc + bias), 1e-5)

outb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
KN.gemm(A, B, bias=bias, act=1, preact=pre, out_bf16=outb)
z = acc + bias
report("epilogue bias+gelu -> bf16", relerr(outb, torch.nn.functional.gelu(z)), 4e-3)
report("epilogue preact bf16", relerr(pre, z), 4e-3)

out = torch.empty(M, N, device=dev)
KN.gemm(A, B, bias=bias, residual=resid, out_f32=out)
report("epilogue bias+residual", relerr(out, acc + bias + resid), 1e-5)

zz = (torch.randn(M, N, device=dev)).to(torch.bfloat16)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
KN.gemm(A, B, dgelu_z=zz, out_bf16=out)
zf = zz.float().requires_grad_(True)
torch.nn.functional.gelu(zf).sum().backward()
report("epilogue dgelu", relerr(out, acc * zf.grad), 4e-3)

out = torch.ones(M, N, device=dev)
KN.gemm(A, B, out_f32=out, accumulate=True, alpha=0.5)
report("epilogue accumulate alpha", relerr(out, 1 + 0.5 * acc), 1e-5)

# ---- element-wise kernels
x = torch.randn(1000, 776, device=dev)
report("cast_bf16", relerr(KN.cast_bf16(x), x.to(torch.bfloat16)), 1e-7)
dst = torch.empty(1000, 776, device=dev, dtype=torch.bfloat16)
cs = torch.zeros(776, device=dev)
KN.cast_colsum(x, dst, cs)
report("cast_colsum cast", relerr(dst, x.to(torch.bfloat16)), 1e-7)
report("cast_colsum sum", relerr(cs, x.sum(0)), 1e-5)
cs2 = torch.zeros(776, device=dev)
KN.colsum_bf16(dst, cs2)
report("colsum_bf16", relerr(cs2, dst.float().sum(0)), 1e-5)
t = KN.transpose_bf16(dst)
report("transpose_bf16", relerr(t, dst.t()), 1e-7)


# ---- LayerNorm forward / backward
for (M, D) in [(1000, 768), (396, 256), (130, 1024)]:
    x = torch.randn(M, D, device=dev) * 2 + 0.5
    gam = torch.randn(D, device=dev)
    bet = torch.randn(D, device=dev)
    yb, yf, mean, rstd = KN.layernorm_fwd(x, gam, bet, 1e-6, out_bf16=True, out_f32=True)
    xr = x.clone().requires_grad_(True)
    gr = gam.clone().requires_grad_(True)
    br = bet.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-6)
    report("ln fwd f32 M=%d D=%d" % (M, D), relerr(yf, ref), 1e-5)
    report("ln fwd bf16 M=%d D=%d" % (M, D), relerr(yb, ref), 4e-3)
    dy = torch.randn(M, D, device=dev)
    resid = torch.randn(M, D, device=dev)
    ref.backward(dy)
    dgam = torch.zeros(D, device=dev)
    dbet = torch.zeros(D, device=dev)
    dx = KN.layernorm_bwd(dy, x, mean, rstd, gam, dgam, dbet, dx_resid=resid)
    report("ln bwd dx (+resid) M=%d D=%d" % (M, D), relerr(dx, xr.grad + resid), 1e-5)
    report("ln bwd dgamma M=%d D=%d" % (M, D), relerr(dgam, gr.grad), 1e-4)
    report("ln bwd dbeta M=%d D=%d" % (M, D), relerr(dbet, br.grad), 1e-4)
    dyb = dy.to(torch.bfloat16)
    dx2 = KN.layernorm_bwd(dyb, x, mean, rstd, gam, torch.zeros_like(gam), torch.zeros_like(gam))
    xr2 = x.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr2, (D,), gam, bet, 1e-6).backward(dyb.float())
    report("ln bwd dx (bf16 dy) M=%d D=%d" % (M, D), relerr(dx2, xr2.grad), 1e-5)


# ---- attention forward / backward (packed projection layouts as the blocks use them)
def attn_case(B, H, Nq, Nk, dh, self_attn):
    D = H * dh
    scale = dh ** -0.5
    if self_attn:
        qkv = rand_bf16(B * Nq, 3 * D)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        q = rand_bf16(B * Nq, D)
        kv = rand_bf16(B * Nk, 2 * D)
        k, v = kv[:, :D], kv[:, D:]
    o, lse = KN.attention_fwd(q, k, v, B, H, Nq, Nk, dh, scale)
    qf = q.float().reshape(B, Nq, H, dh).transpose(1, 2).detach().requires_grad_(True)
    kf = k.float().reshape(B, Nk, H, dh).transpose(1, 2).detach().requires_grad_(True)
    vf = v.float().reshape(B, Nk, H, dh).transpose(1, 2).detach().requires_grad_(True)
    s = (qf @ kf.transpose(-2, -1)) * scale
    ref = torch.softmax(s, -1) @ vf
    ref2 = ref.transpose(1, 2).reshape(B * Nq, D)
    tag = "B=%d H=%d Nq=%d Nk=%d dh=%d" % (B, H, Nq, Nk, dh)
    report("attn fwd o " + tag, relerr(o, ref2), 6e-3)
    report("attn fwd lse " + tag, relerr(lse, torch.logsumexp(s, -1)), 1e-4)
    do = rand_bf16(B * Nq, D)
    ref2.backward(do.float())
    if self_attn:
        dqkv = torch.empty(B * Nq, 3 * D, device=dev, dtype=torch.bfloat16)
        dq, dk, dv = dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:]
    else:
        dq = torch.empty(B * Nq, D, device=dev, dtype=torch.bfloat16)
        dkv = torch.empty(B * Nk, 2 * D, device=dev, dtype=torch.bfloat16)
        dk, dv = dkv[:, :D], dkv[:, D:]
    KN.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, Nq, Nk, dh, scale)
    report("attn bwd dq " + tag, relerr(dq, qf.grad.transpose(1, 2).reshape(B * Nq, D)), 1e-2)
    report("attn bwd dk " + tag, relerr(dk, kf.grad.transpose(1, 2).reshape(B * Nk, D)), 1e-2)
    report("attn bwd dv " + tag, relerr(dv, vf.grad.transpose(1, 2).reshape(B * Nk, D)), 1e-2)


for tc in (0, 1, 3, 7):
    LIB.lib().mmae_attention_set_tc(tc)
    print("== attention tcgen05 path", tc, flush=True)
    for args in [(3, 12, 99, 99, 64, True), (2, 8, 196, 99, 32, False), (2, 8, 196, 196, 32, True),
                 (1, 2, 393, 393, 64, True), (1, 2, 130, 70, 32, False), (2, 1, 17, 5, 64, False),
                 (2, 16, 99, 99, 64, True), (2, 3, 128, 128, 64, True), (1, 2, 100, 33, 64, False)]:
        try:
            attn_case(*args)
        except Exception as e:  # noqa: BLE001
            fails += 1
            print("attn case %s EXC %s" % (str(args), e), flush=True)

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
                                      (768, 768, 12672, 1, 1, 8), (25088, 1024, 256, 0, 0, 1), (25088, 256, 256, 0, 0, 1),
                                      (25088, 256, 1024, 0, 0, 1), (25088, 768, 256, 0, 0, 1), (256, 256, 25088, 1, 1, 16),
                                      (12544, 768, 2048, 0, 0, 1)]:
    A = rand_bf16(M, K_)
    B = rand_bf16(N, K_)
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    fl = 2.0 * M * N * K_
    res = []
    for variant in (0, 1, 2, 3):
        LIB.lib().mmae_gemm_set_variant(variant)
        if split > 1:
            out = torch.zeros(M, N, device=dev)
            ms = time_it(lambda: KN.gemm(Ain, Bin, a_mn=bool(a_mn), b_mn=bool(b_mn), out_f32=out, split_k=split))
        else:
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            ms = time_it(lambda: KN.gemm(Ain, Bin, a_mn=bool(a_mn), b_mn=bool(b_mn), out_bf16=out))
        res.append("v%d %.3f ms (%.0f TF/s)" % (variant, ms, fl / ms / 1e9))
    LIB.lib().mmae_gemm_set_variant(-1)
    ms_ref = time_it(lambda: torch.matmul(A, B.t()))
    print("time M=%d N=%d K=%d a_mn=%d b_mn=%d split=%d : %s | cublas %.3f ms (%.0f TF/s)" %
          (M, N, K_, a_mn, b_mn, split, "  ".join(res), ms_ref, fl / ms_ref / 1e9), flush=True)


# epilogue cost: same GEMM with / without bias, bf16 vs fp32 output
LIB.lib().mmae_gemm_set_variant(-1)
for (M, N, K_) in [(12672, 3072, 768), (12672, 2304, 768), (12672, 768, 768), (25088, 1024, 256), (25088, 256, 256),
                   (25088, 768, 256), (25088, 256, 1024)]:
    A = rand_bf16(M, K_)
    B = rand_bf16(N, K_)
    bias = torch.randn(N, device=dev)
    ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    of = torch.empty(M, N, device=dev)
    fl = 2.0 * M * N * K_
    for tma in (1, 0):
        LIB.lib().mmae_gemm_set_tma_store(tma)
        t0 = time_it(lambda: KN.gemm(A, B, out_bf16=ob))
        t1 = time_it(lambda: KN.gemm(A, B, bias=bias, out_bf16=ob))
        t2 = time_it(lambda: KN.gemm(A, B, bias=bias, out_f32=of))
        t3 = time_it(lambda: torch.nn.functional.linear(A, B, bias.to(torch.bfloat16)))
        print("epilogue tma_store=%d M=%d N=%d K=%d: plain bf16 %.3f ms (%.0f TF/s) | +bias %.3f ms (%.0f) | +bias fp32-out "
              "%.3f ms (%.0f) | cublas+bias %.3f ms (%.0f)" %
              (tma, M, N, K_, t0, fl / t0 / 1e9, t1, fl / t1 / 1e9, t2, fl / t2 / 1e9, t3, fl / t3 / 1e9), flush=True)
LIB.lib().mmae_gemm_set_tma_store(1)

# split-K weight-gradient shapes (fp32 accumulate): TMA reduce-add vs per-lane red.global, over split counts and tiles
for (M, N, K_) in [(256, 256, 25088), (256, 1024, 25088), (1024, 256, 25088), (768, 256, 25088), (768, 768, 12672),
                   (768, 3072, 12672), (3072, 768, 12672), (2304, 768, 12672)]:
    A = rand_bf16(K_, M)
    B = rand_bf16(K_, N)
    of = torch.zeros(M, N, device=dev)
    fl = 2.0 * M * N * K_
    res = []
    for variant in (1, 3, 2):
        LIB.lib().mmae_gemm_set_variant(variant)
        bn = {1: 128, 2: 256, 3: 192}[variant]
        tiles = -(-M // 128) * -(-N // bn)
        for mult in (1, 2):
            split = max(1, (148 * mult) // tiles)
            for tma in (1, 0):
                LIB.lib().mmae_gemm_set_tma_store(tma)
                t = time_it(lambda: KN.gemm(A, B, a_mn=True, b_mn=True, out_f32=of, accumulate=True, split_k=split))
                res.append("bn%d s%d tma%d %.1f us (%.0f)" % (bn, split, tma, t * 1e3, fl / t / 1e9))
    print("wgrad M=%d N=%d K=%d: %s" % (M, N, K_, " | ".join(res)), flush=True)
LIB.lib().mmae_gemm_set_variant(-1)
LIB.lib().mmae_gemm_set_tma_store(1)

# attention timing at the encoder shape
B_, H_, N_, dh_ = 128, 12, 99, 64
qkv = rand_bf16(B_ * N_, 3 * H_ * dh_)
D_ = H_ * dh_
q, k, v = qkv[:, :D_], qkv[:, D_:2 * D_], qkv[:, 2 * D_:]
o = torch.empty(B_ * N_, D_, device=dev, dtype=torch.bfloat16)
fl = 4.0 * B_ * H_ * N_ * N_ * dh_
do = rand_bf16(B_ * N_, D_)
dqkv = torch.empty_like(qkv)
for tc in (0, 1, 3, 7):
    LIB.lib().mmae_attention_set_tc(tc)
    ms = time_it(lambda: KN.attention_fwd(q, k, v, B_, H_, N_, N_, dh_, 0.125, out=o))
    print("time attn fwd enc (tc=%d): %.3f ms (%.1f TF/s useful)" % (tc, ms, fl / ms / 1e9), flush=True)
    o2, lse = KN.attention_fwd(q, k, v, B_, H_, N_, N_, dh_, 0.125)
    ms = time_it(lambda: KN.attention_bwd(q, k, v, o2, do, lse, dqkv[:, :D_], dqkv[:, D_:2 * D_], dqkv[:, 2 * D_:], B_, H_, N_, N_, dh_, 0.125))
    print("time attn bwd enc (tc=%d): %.3f ms (%.1f TF/s useful)" % (tc, ms, 2.5 * fl / ms / 1e9), flush=True)
# decoder attention shapes (warp-MMA kernels)
LIB.lib().mmae_attention_set_tc(3)
for (Nq_, Nk_, self_) in [(196, 196, True), (196, 99, False)]:
    Dd_, Hd_ = 256, 8
    if self_:
        qkv_d = rand_bf16(B_ * Nq_, 3 * Dd_)
        qd, kd, vd = qkv_d[:, :Dd_], qkv_d[:, Dd_:2 * Dd_], qkv_d[:, 2 * Dd_:]
    else:
        qd = rand_bf16(B_ * Nq_, Dd_)
        kvd = rand_bf16(B_ * Nk_, 2 * Dd_)
        kd, vd = kvd[:, :Dd_], kvd[:, Dd_:]
    for tc_ in (3, 0, 7):
      LIB.lib().mmae_attention_set_tc(tc_)
      od, lsed = KN.attention_fwd(qd, kd, vd, B_, Hd_, Nq_, Nk_, 32, 32 ** -0.5)
      ms = time_it(lambda: KN.attention_fwd(qd, kd, vd, B_, Hd_, Nq_, Nk_, 32, 32 ** -0.5, out=od))
      fld = 4.0 * B_ * Hd_ * Nq_ * Nk_ * 32
      dod = rand_bf16(B_ * Nq_, Dd_)
      dq_, dk_, dv_ = torch.empty_like(qd), torch.empty_like(kd.contiguous()), torch.empty_like(vd.contiguous())
      msb = time_it(lambda: KN.attention_bwd(qd, kd, vd, od, dod, lsed, dq_, dk_, dv_, B_, Hd_, Nq_, Nk_, 32, 32 ** -0.5))
      print("time attn dec tc=%d %dx%d: fwd %.3f ms (%.1f TF/s)  bwd %.3f ms (%.1f TF/s)" % ((tc_,) +
            (Nq_, Nk_, ms, fld / ms / 1e9, msb, 2.5 * fld / msb / 1e9)), flush=True)
LIB.lib().mmae_attention_set_tc(3)
x = torch.randn(12672, 768, device=dev)
gam = torch.ones(768, device=dev)
bet = torch.zeros(768, device=dev)
ms = time_it(lambda: KN.layernorm_fwd(x, gam, bet))
print("time ln fwd 12672x768: %.3f ms (%.0f GB/s)" % (ms, 12672 * 768 * 6 / ms / 1e6), flush=True)

print("FAILS", fails)
sys.exit(1 if fails else 0)
