"""GPU: phase timeline of CTA 0 of the warp-specialised attention kernels (clock64 stamps, attention_ws.cu trace_stamp).

    python scripts/gpu_trace_attention_ws.py            # prints per-item phase offsets in SM clock cycles
stamps: 0 loads issued, 1 loads landed (MMA warp), 2 S issued, 3 PV / grads issued, 4 compute waits for S, 5 S ready,
        6 row max known (fwd), 7 P (/dS) written, 8 O / grads ready, 9 epilogue done"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimae_b200 import _lib as L  # noqa: E402
from multimae_b200 import kernels as KN  # noqa: E402

dev = torch.device("cuda")
lib = L.lib()
DEFAULT = int(os.environ.get("MMAE_ATTN_TC", "3"))


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)


def show(title, tr, n=12):
    t = tr.cpu().reshape(64, 16)
    base = int(t[0, 0]) if int(t[0, 0]) else int(t[t > 0].min())
    print(title)
    print("  it |" + "".join("%9s" % s for s in ("ld_iss", "ld_land", "S_iss", "PV_iss", "c_wait", "S_rdy", "max", "P_done", "O_rdy", "ep_done")))
    for it in range(n):
        row = t[it]
        if int(row[4]) == 0:
            break
        print("  %2d |" % it + "".join("%9d" % (int(v) - base if int(v) else -1) for v in row[:10]))


for name, (B, H, Nq, Nk, dh, sa) in (("enc 99x99x64", (128, 12, 99, 99, 64, True)), ("dec 196x196x32", (128, 8, 196, 196, 32, True))):
    D, scale = H * dh, dh ** -0.5
    if sa:
        qkv = bf(B * Nq, 3 * D)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        q, kv = bf(B * Nq, D), bf(B * Nk, 2 * D)
        k, v = kv[:, :D], kv[:, D:]
    o = torch.empty(B * Nq, D, device=dev, dtype=torch.bfloat16)
    do = bf(B * Nq, D)
    dq, dk, dv = torch.empty_like(o), torch.empty(B * Nk, D, device=dev, dtype=torch.bfloat16), torch.empty(B * Nk, D, device=dev, dtype=torch.bfloat16)
    lib.mmae_attention_set_tc(DEFAULT | 32 | 64)
    _, lse = KN.attention_fwd(q, k, v, B, H, Nq, Nk, dh, scale, out=o)
    KN.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, Nq, Nk, dh, scale)
    torch.cuda.synchronize()
    tr = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
    lib.mmae_attention_ws_set_trace(tr.data_ptr())
    KN.attention_fwd(q, k, v, B, H, Nq, Nk, dh, scale, out=o)
    torch.cuda.synchronize()
    show("forward  %s (cycles since CTA 0's first load)" % name, tr)
    tr.zero_()
    KN.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, Nq, Nk, dh, scale)
    torch.cuda.synchronize()
    show("backward %s" % name, tr)
    lib.mmae_attention_ws_set_trace(None)
lib.mmae_attention_set_tc(DEFAULT)
