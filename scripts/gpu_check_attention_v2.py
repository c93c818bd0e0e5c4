"""GPU: validate and time the EXPERIMENTAL persistent encoder-attention forward (attn_tc_fwd_persistent_kernel, bit 4 of
mmae_attention_set_tc; written after round 1's GPU budget was spent) against attn_tc_fwd_kernel and an fp32 reference.

    timeout 120 python scripts/gpu_check_attention_v2.py          # bounded: a wrong barrier phase would hang, not fail

The two kernels do the same arithmetic in the same order per (batch, head), so their outputs must be bit-identical; the
fp32 softmax-attention reference bounds both (1e-2 relative L2, like tests/test_cuda_kernels.py).  Exit code 0 only if every
case passes.  Make it the default (attention.cu dispatch) only if it also wins the timing at 128 x 12 x 99 x 99 x 64."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimae_b200 import _lib as L  # noqa: E402
from multimae_b200 import kernels as KN  # noqa: E402

dev = torch.device("cuda")
lib = L.lib()


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def time_us(fn, iters=30):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
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


def main():
    torch.manual_seed(0)
    fails = 0
    # (B, H, N) at head_dim 64 with N <= 128: the shapes attn_tc_supported() accepts; 1 and 2 items per CTA, 600+ items,
    # fewer items than CTAs, one full 128-row tile, a short sequence
    for B, H, N in ((3, 12, 99), (128, 12, 99), (64, 16, 99), (1, 2, 99), (2, 3, 128), (5, 2, 17), (37, 12, 99)):
        D, scale = H * 64, 64 ** -0.5
        qkv = (torch.randn(B * N, 3 * D, device=dev) * 0.5).to(torch.bfloat16)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        lib.mmae_attention_set_tc(3)
        o1, lse1 = KN.attention_fwd(q, k, v, B, H, N, N, 64, scale)
        lib.mmae_attention_set_tc(3 | 16)
        o2, lse2 = KN.attention_fwd(q, k, v, B, H, N, N, 64, scale)
        torch.cuda.synchronize()
        qf = q.float().reshape(B, N, H, 64).transpose(1, 2)
        kf = k.float().reshape(B, N, H, 64).transpose(1, 2)
        vf = v.float().reshape(B, N, H, 64).transpose(1, 2)
        s = (qf @ kf.transpose(-2, -1)) * scale
        ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B * N, D)
        same = torch.equal(o1, o2) and torch.equal(lse1, lse2)
        err = rel_l2(o2, ref)
        err_lse = rel_l2(lse2, torch.logsumexp(s, -1))
        ok = same and err < 1e-2 and err_lse < 1e-4
        fails += 0 if ok else 1
        print("B=%3d H=%2d N=%3d  bit-identical to the validated kernel: %s  rel-l2 vs fp32 %.2e  lse %.2e  %s"
              % (B, H, N, same, err, err_lse, "ok" if ok else "FAIL"), flush=True)
    B, H, N, D = 128, 12, 99, 768
    qkv = (torch.randn(B * N, 3 * D, device=dev) * 0.5).to(torch.bfloat16)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    out = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
    for flags, name in ((3, "attn_tc_fwd_kernel (default)"), (3 | 16, "attn_tc_fwd_persistent_kernel")):
        lib.mmae_attention_set_tc(flags)
        us = time_us(lambda: KN.attention_fwd(q, k, v, B, H, N, N, 64, 0.125, out=out))
        print("time %-34s 128x12x99x99x64: %.1f us (HBM floor 12 us: Q/K/V in + O out = 78 MB)" % (name, us), flush=True)
    lib.mmae_attention_set_tc(3)
    print("FAILS %d" % fails)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
