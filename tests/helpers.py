"""Shared test helpers (CPU + GPU tests, and tests/golden/make_golden.py)."""
import math

import torch


def formula_fill_(named_params):
    """Deterministic, RNG-free parameter values so large fixtures need not store weights.

    `named_params`: iterable of (name, tensor) — iterated in sorted-name order; frozen pos_emb tables are skipped."""
    with torch.no_grad():
        todo = sorted([(n, p) for n, p in named_params if not n.endswith("pos_emb")], key=lambda kv: kv[0])
        for i, (name, p) in enumerate(todo):
            idx = torch.arange(p.numel(), dtype=torch.float64)
            wave = torch.sin(idx * 0.37 + i * 1.3) + 0.5 * torch.cos(idx * 0.011 + i)
            if p.dim() >= 2:
                fan_in = p[0].numel()
                vals = wave * (0.7 / math.sqrt(fan_in))
            elif name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("_norm.weight"):
                vals = 1.0 + 0.1 * wave
            else:
                vals = 0.05 * wave
            p.copy_(vals.reshape(p.shape).to(p.dtype))


def digest(t, n=256):
    """(l2 norm, n strided samples) of a tensor — compact stand-in for a full gradient in fixtures."""
    flat = t.detach().float().flatten()
    step = max(1, flat.numel() // n)
    return {"norm": flat.norm().clone(), "samples": flat[::step][:n].clone(), "step": step}


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))
