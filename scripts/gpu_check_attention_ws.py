"""GPU: validate and time the warp-specialised persistent tcgen05 attention kernels (attention_ws.cu; bits 5 / 6 of
mmae_attention_set_tc) against an fp32 torch reference and the current default kernels.

    timeout 300 python scripts/gpu_check_attention_ws.py [fwd|bwd|all]        # a barrier-protocol bug shows up as a hang
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimae_b200 import _lib as L  # noqa: E402
from multimae_b200 import kernels as KN  # noqa: E402

dev = torch.device("cuda")
lib = L.lib()
what = sys.argv[1] if len(sys.argv) > 1 else "all"
DEFAULT = int(os.environ.get("MMAE_ATTN_TC", "3"))


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)


def time_us(fn, iters=20, cold=True):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3


CASES = [(3, 12, 99, 99, 64, True), (2, 8, 196, 99, 32, False), (2, 8, 196, 196, 32, True), (1, 2, 130, 70, 32, False),
         (2, 1, 17, 5, 64, False), (2, 16, 99, 99, 64, True), (2, 3, 128, 128, 64, True), (1, 2, 100, 33, 64, False),
         (1, 2, 256, 256, 64, True), (1, 4, 200, 129, 32, False), (37, 12, 99, 99, 64, True), (40, 8, 196, 196, 32, True),
         (1, 2, 393, 200, 64, False), (128, 12, 99, 99, 64, True)]


def make(B, H, Nq, Nk, dh, self_attn):
    D = H * dh
    if self_attn:
        qkv = bf(B * Nq, 3 * D)
        return qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    q, kv = bf(B * Nq, D), bf(B * Nk, 2 * D)
    return q, kv[:, :D], kv[:, D:]


def main():
    torch.manual_seed(0)
    fails = 0
    for case in CASES:
        B, H, Nq, Nk, dh, self_attn = case
        D, scale = H * dh, dh ** -0.5
        q, k, v = make(*case)
        qf = q.float().reshape(B, Nq, H, dh).transpose(1, 2).detach().requires_grad_(True)
        kf = k.float().reshape(B, Nk, H, dh).transpose(1, 2).detach().requires_grad_(True)
        vf = v.float().reshape(B, Nk, H, dh).transpose(1, 2).detach().requires_grad_(True)
        s = (qf @ kf.transpose(-2, -1)) * scale
        ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B * Nq, D)
        do = bf(B * Nq, D)
        ref.backward(do.float())
        unt = lambda t, N: t.transpose(1, 2).reshape(B * N, D)  # noqa: E731
        msg = "B=%3d H=%2d %3dx%3d dh=%d" % (B, H, Nq, Nk, dh)
        if what in ("fwd", "all") and Nk <= 256:
            lib.mmae_attention_set_tc(DEFAULT | 32)
            o, lse = KN.attention_fwd(q, k, v, B, H, Nq, Nk, dh, scale)
            torch.cuda.synchronize()
            e1, e2 = rel_l2(o, ref), rel_l2(lse, torch.logsumexp(s, -1))
            ok = e1 < 1e-2 and e2 < 1e-4
            fails += 0 if ok else 1
            msg += "  fwd o %.2e lse %.2e %s" % (e1, e2, "ok" if ok else "FAIL")
        if what in ("bwd", "all") and Nq <= 256:
            lib.mmae_attention_set_tc(DEFAULT)                       # forward state from the validated kernels
            o, lse = KN.attention_fwd(q, k, v, B, H, Nq, Nk, dh, scale)
            dq, dk, dv = torch.empty(B * Nq, D, device=dev, dtype=torch.bfloat16), torch.empty(B * Nk, D, device=dev, dtype=torch.bfloat16), torch.empty(B * Nk, D, device=dev, dtype=torch.bfloat16)
            lib.mmae_attention_set_tc(DEFAULT | 64)
            KN.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, Nq, Nk, dh, scale)
            torch.cuda.synchronize()
            errs = [rel_l2(dq, unt(qf.grad, Nq)), rel_l2(dk, unt(kf.grad, Nk)), rel_l2(dv, unt(vf.grad, Nk))]
            ok = max(errs) < 2e-2
            fails += 0 if ok else 1
            msg += "  bwd dq %.2e dk %.2e dv %.2e %s" % (errs[0], errs[1], errs[2], "ok" if ok else "FAIL")
        print(msg, flush=True)
    # ---------------------------------------------------------------- timing at the MultiMAE-B bs=128 shapes
    # back-to-back launches between ONE pair of events (the host-side tensor-map encoding of a call hides behind the previous
    # kernel); "cold": the operand sets rotate through > 2 x L2 of distinct buffers, "warm": one set, L2-resident
    def timed_loop(fn_of_set, nsets, iters=24):
        for i in range(3):
            fn_of_set(i % nsets)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            fn_of_set(i % nsets)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e3

    for name, (B, H, Nq, Nk, dh, self_attn) in (("enc 99x99x64", (128, 12, 99, 99, 64, True)),
                                                ("dec 196x196x32", (128, 8, 196, 196, 32, True)),
                                                ("dec 196x99x32", (128, 8, 196, 99, 32, False))):
        D, scale = H * dh, dh ** -0.5
        NS = 4
        sets = []
        for _ in range(NS):
            q, k, v = make(B, H, Nq, Nk, dh, self_attn)
            o = torch.empty(B * Nq, D, device=dev, dtype=torch.bfloat16)
            do = bf(B * Nq, D)
            dq, dk, dv = torch.empty_like(o), torch.empty(B * Nk, D, device=dev, dtype=torch.bfloat16), torch.empty(B * Nk, D, device=dev, dtype=torch.bfloat16)
            lib.mmae_attention_set_tc(DEFAULT)
            _, lse = KN.attention_fwd(q, k, v, B, H, Nq, Nk, dh, scale, out=o)
            sets.append((q, k, v, o, do, lse, dq, dk, dv))
        for flags, tag in ((DEFAULT, "default"), (DEFAULT | 32 | 64, "ws")):
            lib.mmae_attention_set_tc(flags)
            line = "time %-15s %-8s" % (name, tag)
            for nsets, label in ((NS, "cold"), (1, "warm")):
                if what in ("fwd", "all"):
                    def f(i):
                        q, k, v, o, do, lse, dq, dk, dv = sets[i]
                        KN.attention_fwd(q, k, v, B, H, Nq, Nk, dh, scale, out=o)
                    line += "  fwd %s %6.1f us" % (label, timed_loop(f, nsets))
                if what in ("bwd", "all"):
                    def g(i):
                        q, k, v, o, do, lse, dq, dk, dv = sets[i]
                        KN.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, Nq, Nk, dh, scale)
                    line += "  bwd(+delta) %s %6.1f us" % (label, timed_loop(g, nsets))
            print(line, flush=True)
    lib.mmae_attention_set_tc(DEFAULT)
    print("FAILS %d" % fails)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
