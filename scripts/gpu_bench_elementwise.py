"""GPU: achieved HBM bandwidth of the memory-bound kernels of the step at their real shapes.

Each kernel runs over R rotating buffer sets whose total exceeds the 126 MB L2, so the figures are DRAM figures; bytes
are the algorithmic bytes (every operand read once, every result written once).

    python scripts/gpu_bench_elementwise.py > profiles/rNN_elementwise_bandwidth.txt
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from multimae_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.lib()
PEAK = 6575.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def st():
    return L.current_stream()


def time_sets(fn, nsets, iters=40):
    for i in range(nsets):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nsets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, shape, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print("%-34s %-16s %8.1f MB %8.1f us %8.0f GB/s  %5.1f%% of %.0f" %
          (name, "x".join(map(str, shape)), nbytes / 1e6, ms * 1e3, gbs, 100 * gbs / PEAK, PEAK), flush=True)


def nsets_for(nbytes):
    return max(2, int(400e6 // nbytes) + 1)


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)


def f32(*s):
    return torch.randn(*s, device=dev)


for (M, N) in [(12672, 3072), (25088, 1024)]:
    nb = M * N * 2
    R = nsets_for(2 * nb)
    z, h = [bf(M, N) for _ in range(R)], [bf(M, N) for _ in range(R)]
    report("gelu fwd (z -> h)", (M, N), 2 * nb,
           time_sets(lambda i: L.check(lib.mmae_gelu_bf16(z[i].data_ptr(), h[i].data_ptr(), M * N, 0, st())), R))
    cs = torch.zeros(N, device=dev)
    report("dgelu*dh + colsum (in place)", (M, N), 3 * nb,
           time_sets(lambda i: L.check(lib.mmae_dgelu_colsum_bf16(z[i].data_ptr(), h[i].data_ptr(), N, cs.data_ptr(), M, N, st())), R))
    del z, h

for (M, D) in [(12672, 768), (25088, 256)]:
    R = nsets_for(M * D * 16)
    x, add, xs, y = [f32(M, D) for _ in range(R)], [bf(M, D) for _ in range(R)], [f32(M, D) for _ in range(R)], [bf(M, D) for _ in range(R)]
    gam, bet = f32(D), f32(D)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    report("layernorm fwd", (M, D), M * D * 6,
           time_sets(lambda i: L.check(lib.mmae_layernorm_forward(x[i].data_ptr(), D, gam.data_ptr(), bet.data_ptr(), y[i].data_ptr(), D,
                                                                  None, 0, mean.data_ptr(), rstd.data_ptr(), M, D, 1e-6, st())), R))
    report("add + layernorm fwd", (M, D), M * D * 12,
           time_sets(lambda i: L.check(lib.mmae_add_layernorm_forward(x[i].data_ptr(), D, add[i].data_ptr(), D, xs[i].data_ptr(), D,
                                                                      gam.data_ptr(), bet.data_ptr(), y[i].data_ptr(), D,
                                                                      mean.data_ptr(), rstd.data_ptr(), M, D, 1e-6, st())), R))
    dg, db, cs = torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    dx = [f32(M, D) for _ in range(R)]
    report("layernorm bwd (+resid)", (M, D), M * D * (2 + 4 + 4 + 4),
           time_sets(lambda i: L.check(lib.mmae_layernorm_backward(add[i].data_ptr(), 1, D, x[i].data_ptr(), D, mean.data_ptr(),
                                                                   rstd.data_ptr(), gam.data_ptr(), xs[i].data_ptr(), D,
                                                                   dx[i].data_ptr(), D, dg.data_ptr(), db.data_ptr(), M, D, st())), R))
    report("layernorm bwd ex (+resid,+bf16)", (M, D), M * D * (2 + 4 + 4 + 4 + 2),
           time_sets(lambda i: L.check(lib.mmae_layernorm_backward_ex(add[i].data_ptr(), 1, D, x[i].data_ptr(), D, mean.data_ptr(),
                                                                      rstd.data_ptr(), gam.data_ptr(), xs[i].data_ptr(), D,
                                                                      dx[i].data_ptr(), D, dg.data_ptr(), db.data_ptr(),
                                                                      y[i].data_ptr(), D, cs.data_ptr(), M, D, st())), R))
    report("cast f32->bf16 + colsum", (M, D), M * D * 6,
           time_sets(lambda i: L.check(lib.mmae_cast_colsum_f32(x[i].data_ptr(), D, y[i].data_ptr(), D, cs.data_ptr(), M, D, st())), R))
    report("colsum bf16", (M, D), M * D * 2,
           time_sets(lambda i: L.check(lib.mmae_colsum_bf16(y[i].data_ptr(), D, cs.data_ptr(), M, D, st())), R))
    report("add bf16 onto f32", (M, D), M * D * 10,
           time_sets(lambda i: L.check(lib.mmae_add_bf16_f32(x[i].data_ptr(), add[i].data_ptr(), xs[i].data_ptr(), M * D, st())), R))
    del x, add, xs, y, dx

(M, D) = (12672, 2304)
R = nsets_for(M * D * 2)
y = [bf(M, D) for _ in range(R)]
cs = torch.zeros(D, device=dev)
report("colsum bf16", (M, D), M * D * 2,
       time_sets(lambda i: L.check(lib.mmae_colsum_bf16(y[i].data_ptr(), D, cs.data_ptr(), M, D, st())), R))
del y

n = 111_000_000   # ~ MultiMAE-B parameter count
p, g, m, v = f32(n), f32(n) * 1e-3, torch.zeros(n, device=dev), torch.zeros(n, device=dev)
report("adamw (flat)", (n,), n * 28,
       time_sets(lambda i: L.check(lib.mmae_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-4, 0.9, 0.95,
                                                       1e-8, 0.05, 1, None, None, st())), 1, iters=10))
out2, nrm = torch.zeros(2, device=dev), torch.zeros(1, device=dev)
report("grad unscale + norm", (n,), n * 4,
       time_sets(lambda i: L.check(lib.mmae_grad_unscale_norm(g.data_ptr(), n, None, 1.0, 1.0, out2.data_ptr(), nrm.data_ptr(), st())), 1, iters=10))
w = f32(n)
wb = torch.empty(n, device=dev, dtype=torch.bfloat16)
report("cast f32->bf16 (flat)", (n,), n * 6,
       time_sets(lambda i: L.check(lib.mmae_cast_f32_to_bf16(w.data_ptr(), wb.data_ptr(), n, st())), 1, iters=10))
