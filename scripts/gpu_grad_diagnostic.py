"""Per-parameter gradient error table: CUDA path vs fp32 oracle, next to stock torch bf16-autocast of the same oracle
functions on the GPU (the noise level the reference's own mixed-precision path would show).  Run under gpurun."""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from helpers import formula_fill_, rel_l2  # noqa: E402
from oracle import multimae_oracle as O  # noqa: E402
from test_cuda_parity import _build_model, _load, _oracle_cfg, _run_cuda_step  # noqa: E402

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "cuda_small.pt"
fx = _load("tests/golden", name)
c = fx["config"]
model = _build_model(c)
formula_fill_(list(model.named_parameters()))
model = model.to(dev).train()
triple = ({k: v.to(dev) for k, v in fx["task_masks"].items()}, fx["ids_keep"].to(dev), fx["ids_restore"].to(dev))
print("running cuda step", flush=True)
preds, masks, losses = _run_cuda_step(model, fx["inputs"], triple, dev)
print("cuda step done", flush=True)


def oracle_grads(device, autocast):
    cfg = _oracle_cfg(c)
    p = {k: v.to(device) for k, v in O.init_params(cfg).items()}
    train = O.trainable(p)
    formula_fill_(list(train.items()))
    for v in train.values():
        v.requires_grad_(True)
    x = {k: v.to(device) for k, v in fx["inputs"].items()}
    tm = {k: v.to(device) for k, v in fx["task_masks"].items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        ls, pr = O.step_losses(p, x, cfg, tm, fx["ids_keep"].to(device), fx["ids_restore"].to(device))
    sum(ls.values()).backward()
    return {k: v.grad.detach().float().cpu() for k, v in train.items()}, {k: v.detach().float().cpu() for k, v in pr.items()}


ref, ref_preds = oracle_grads(torch.device("cpu"), False)
print("cpu oracle done", flush=True)
amp, amp_preds = oracle_grads(dev, True)
print("gpu amp oracle done", flush=True)
named = dict(model.named_parameters())
rows = []
for k, g in ref.items():
    rows.append((rel_l2(named[k].grad, g), rel_l2(amp[k], g), float(g.norm()), k))
rows.sort(reverse=True)
print("%-10s %-10s %-10s %s" % ("ours", "torch-amp", "|g|", "param"))
for r in rows[:40]:
    print("%-10.4f %-10.4f %-10.3e %s" % r)
allo = torch.cat([named[k].grad.detach().float().cpu().flatten() for k in ref])
allr = torch.cat([ref[k].flatten() for k in ref])
alla = torch.cat([amp[k].flatten() for k in ref])
print("GLOBAL rel-l2 ours %.4f torch-amp %.4f ; median ours %.4f amp %.4f" %
      (rel_l2(allo, allr), rel_l2(alla, allr), sorted(r[0] for r in rows)[len(rows) // 2], sorted(r[1] for r in rows)[len(rows) // 2]))
for k in ref_preds:
    print("pred %-9s ours %.4f amp %.4f" % (k, rel_l2(preds[k], ref_preds[k]), rel_l2(amp_preds[k], ref_preds[k])))
