"""GPU: one launch of each attention kernel variant at its MultiMAE-B (bs 128) shapes between cudaProfilerStart/Stop, for

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_attn \
        python scripts/gpu_ncu_attention.py [warm]

`warm` leaves the operands in L2 (as in the training step, where the QKV GEMM has just written them); the default flushes
the L2 before each profiled launch.  MMAE_ATTN_TC_LIST (comma separated switch values) selects the variants."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from multimae_b200 import _lib as L  # noqa: E402
from multimae_b200 import kernels as KN  # noqa: E402

dev = torch.device("cuda:0")
lib = L.lib()
lib.mmae_set_pdl(0)
warm = len(sys.argv) > 1 and sys.argv[1] == "warm"
variants = [int(v) for v in os.environ.get("MMAE_ATTN_TC_LIST", "3").split(",")]
shapes = os.environ.get("MMAE_ATTN_SHAPES", "enc,dec196,dec99").split(",")


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)


flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def cold():
    if not warm:
        flush.zero_()
    torch.cuda.synchronize()


SHAPES = {"enc": (128, 12, 99, 99, 64), "dec196": (128, 8, 196, 196, 32), "dec99": (128, 8, 196, 99, 32)}
jobs = []
for name in shapes:
    B_, H_, Nq, Nk, dh_ = SHAPES[name]
    Dm = H_ * dh_
    q = bf(B_ * Nq, Dm)
    kv = bf(B_ * Nk, 2 * Dm)
    k, v = kv[:, :Dm], kv[:, Dm:]
    o = torch.empty(B_ * Nq, Dm, device=dev, dtype=torch.bfloat16)
    do = bf(B_ * Nq, Dm)
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    scale = dh_ ** -0.5
    for tc in variants:
        state = {}

        def fwd(q=q, k=k, v=v, o=o, B_=B_, H_=H_, Nq=Nq, Nk=Nk, dh_=dh_, scale=scale, state=state, tc=tc):
            lib.mmae_attention_set_tc(tc)
            state["lse"] = KN.attention_fwd(q, k, v, B_, H_, Nq, Nk, dh_, scale, out=o)[1]

        def bwd(q=q, k=k, v=v, o=o, do=do, dq=dq, dkv=dkv, Dm=Dm, B_=B_, H_=H_, Nq=Nq, Nk=Nk, dh_=dh_, scale=scale,
                state=state, tc=tc):
            lib.mmae_attention_set_tc(tc)
            KN.attention_bwd(q, k, v, o, do, state["lse"], dq, dkv[:, :Dm], dkv[:, Dm:], B_, H_, Nq, Nk, dh_, scale)

        jobs.append(("%s tc=%d fwd" % (name, tc), fwd))
        jobs.append(("%s tc=%d bwd" % (name, tc), bwd))

for name, fn in jobs:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fn in jobs:
    cold()
    fn()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
lib.mmae_attention_set_tc(3)
print("profiled:", [n for n, _ in jobs])
