"""CPU oracle of the MultiMAE pre-training hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A plain-PyTorch fp32 functional restatement of the reference algorithm (EPFL-VILAB/MultiMAE @ 66910f5).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this file;
nothing under `multimae_b200/` does.  It is written against a flat `params` dict that uses the reference's
state_dict key names, so a reference checkpoint / `state_dict()` can be fed in unchanged.

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by the reference itself).
This restatement is pinned instead against the live reference imported from /root/reference in the authoring
container: `tests/golden/make_golden.py` records reference outputs (preds, masks, losses, gradients) as fixtures under
`tests/golden/`, and `tests/test_oracle_golden.py` checks this file against them.

Every function cites the reference lines (path:line under /root/reference) it restates.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------------------------
@dataclass
class DomainSpec:
    """One modality. kind 'image': dense channels; kind 'semseg': class-id map embedded to dim_class_emb."""
    name: str
    kind: str               # 'image' | 'semseg'
    channels: int           # image: input channels; semseg: number of classes
    stride_level: int = 1   # run_pretraining_multimae.py:49-72 (DOMAIN_CONF)
    dim_class_emb: int = 64

    def patch(self, patch_size_full: int) -> int:
        return max(1, patch_size_full // self.stride_level)   # multimae/input_adapters.py:61-62


@dataclass
class OracleConfig:
    in_domains: List[DomainSpec]
    out_tasks: List[Tuple[str, DomainSpec]]   # (output-adapter key, spec of the domain it reconstructs)
    dim: int = 768
    depth: int = 12
    heads: int = 12
    dec_dim: int = 256
    dec_depth: int = 2
    dec_heads: int = 8
    patch_size: int = 16
    num_global_tokens: int = 1
    eps: float = 1e-6
    posemb_grid: int = 14   # image_size 224 // 16 (adapters are always built with image_size=224)
    use_task_queries: bool = True               # --decoder_use_task_queries (run_pretraining_multimae.py:264,278)
    context_tasks: Optional[List[str]] = None   # names with a task embedding in every decoder; None: the in_domains


RGB = DomainSpec("rgb", "image", 3, 1)
DEPTH = DomainSpec("depth", "image", 1, 1)
SEMSEG = DomainSpec("semseg", "semseg", 133, 4)
DOMAINS = {"rgb": RGB, "depth": DEPTH, "semseg": SEMSEG}


def make_config(in_domains=("rgb", "depth", "semseg"), out_domains=None, extra_norm_pix=True, size="base"):
    """Mirror of get_model()'s wiring (run_pretraining_multimae.py:243-293)."""
    out_domains = list(in_domains) if out_domains is None else list(out_domains)
    outs = [(d, DOMAINS[d]) for d in out_domains]
    if extra_norm_pix:
        outs.append(("norm_rgb", RGB))
    kw = dict(dim=768, depth=12, heads=12) if size == "base" else dict(dim=1024, depth=24, heads=16)
    return OracleConfig(in_domains=[DOMAINS[d] for d in in_domains], out_tasks=outs, **kw)


# --------------------------------------------------------------------------------------------------------------
# positional embedding (multimae/multimae_utils.py:29-45)
# --------------------------------------------------------------------------------------------------------------
def sincos_posemb(h: int, w: int, dim: int, temperature: float = 10000.0) -> torch.Tensor:
    """[1, dim, h, w] table.  NOTE the reference's meshgrid(w, h) 'ij' ordering followed by a '(h w)' reshape."""
    assert dim % 4 == 0
    gw, gh = torch.meshgrid(torch.arange(w, dtype=torch.float32), torch.arange(h, dtype=torch.float32), indexing="ij")
    quarter = dim // 4
    omega = 1.0 / (temperature ** (torch.arange(quarter, dtype=torch.float32) / quarter))
    ow = gw.flatten()[:, None] * omega[None, :]
    oh = gh.flatten()[:, None] * omega[None, :]
    table = torch.cat([ow.sin(), ow.cos(), oh.sin(), oh.cos()], dim=1)       # [(w*h), dim]
    return table.reshape(h, w, dim).permute(2, 0, 1).unsqueeze(0).contiguous()


def resized_posemb(table: torch.Tensor, nh: int, nw: int, mode: str) -> torch.Tensor:
    """[nh*nw, dim] rows.  input adapters: bicubic (image) / bilinear (semseg) — multimae/input_adapters.py:113,235;
    output adapters: bilinear — multimae/output_adapters.py:172."""
    kw = dict(align_corners=False) if mode in ("bicubic", "bilinear") else {}
    t = F.interpolate(table, size=(nh, nw), mode=mode, **kw)
    return t.flatten(2).transpose(1, 2)[0]


# --------------------------------------------------------------------------------------------------------------
# mask sampling (multimae/multimae.py:164-218) as a pure function of the random draws
# --------------------------------------------------------------------------------------------------------------
def sample_masks(shares: torch.Tensor, noises: List[torch.Tensor], noise_all: torch.Tensor, num_encoded: int):
    """shares [B,T] (Dirichlet draw), noises[t] [B,N_t] and noise_all [B,sum N_t] uniform in [0,1).

    Returns (task_masks list of [B,N_t] int64, ids_keep [B,num_encoded], ids_restore [B,sum N_t]).
    Ties are broken by lower index first (stable sort); the reference's argsort is unstable, so compare on
    tie-free draws."""
    per_task = (shares * num_encoded).round().long()                       # :189  (round half to even)
    masks = []
    for t, noise in enumerate(noises):
        order = torch.argsort(noise, dim=1, stable=True)                   # :196
        # :197-200 — position j is visible iff the INDEX of the j-th smallest noise is < k_t
        masks.append((order >= per_task[:, t:t + 1]).long())
    mask_all = torch.cat(masks, dim=1)
    ids_shuffle = torch.argsort(mask_all.float() + noise_all, dim=1, stable=True)   # :204
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)           # :205
    ids_keep = ids_shuffle[:, :num_encoded]                                # :206
    final = torch.ones_like(mask_all)                                      # :209-212
    final[:, :num_encoded] = 0
    final = torch.gather(final, 1, ids_restore)
    sizes = [n.shape[1] for n in noises]
    return list(torch.split(final, sizes, dim=1)), ids_keep, ids_restore   # :214


# --------------------------------------------------------------------------------------------------------------
# building blocks (multimae/multimae_utils.py:138-232)
# --------------------------------------------------------------------------------------------------------------
def _ln(x, p, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), p[prefix + ".weight"], p[prefix + ".bias"], eps)


def _lin(x, p, prefix):
    return x @ p[prefix + ".weight"].t() + p[prefix + ".bias"]


def _mlp(x, p, prefix):                                                     # :148-155, exact-erf GELU
    return _lin(F.gelu(_lin(x, p, prefix + ".fc1")), p, prefix + ".fc2")


def _heads(t, h):                                                           # [B,N,h*dh] -> [B,h,N,dh]
    b, n, c = t.shape
    return t.reshape(b, n, h, c // h).transpose(1, 2)


def _attend(q, k, v, heads):                                                # :175-179 / :207-211
    q, k, v = _heads(q, heads), _heads(k, heads), _heads(v, heads)
    scale = q.shape[-1] ** -0.5
    w = torch.softmax((q @ k.transpose(-2, -1)) * scale, dim=-1)
    o = w @ v
    return o.transpose(1, 2).reshape(o.shape[0], o.shape[2], -1)


def _self_attention(x, p, prefix, heads):                                   # :170-182
    q, k, v = _lin(x, p, prefix + ".qkv").chunk(3, dim=-1)                  # rows ordered [q;k;v], head-major
    return _lin(_attend(q, k, v, heads), p, prefix + ".proj")


def _cross_attention(x, ctx, p, prefix, heads):                             # :199-214
    q = _lin(x, p, prefix + ".q")
    k, v = _lin(ctx, p, prefix + ".kv").chunk(2, dim=-1)
    return _lin(_attend(q, k, v, heads), p, prefix + ".proj")


def _block(x, p, prefix, heads, eps):                                       # :229-232 (drop_path = 0)
    x = x + _self_attention(_ln(x, p, prefix + ".norm1", eps), p, prefix + ".attn", heads)
    return x + _mlp(_ln(x, p, prefix + ".norm2", eps), p, prefix + ".mlp")


# --------------------------------------------------------------------------------------------------------------
# input adapters (multimae/input_adapters.py:97-119, 215-241)
# --------------------------------------------------------------------------------------------------------------
def _patchify_linear(x, weight, bias, P):
    """Conv2d(kernel=stride=P) written as unfold + matmul.  x [B,C,H,W] -> [B, nh*nw, D]."""
    B, C, H, W = x.shape
    nh, nw = H // P, W // P
    cols = x.reshape(B, C, nh, P, nw, P).permute(0, 2, 4, 1, 3, 5).reshape(B, nh * nw, C * P * P)
    return cols @ weight.reshape(weight.shape[0], -1).t() + bias


def embed_domain(x, p, spec: DomainSpec, cfg: OracleConfig):
    prefix = "input_adapters.%s" % spec.name
    P = spec.patch(cfg.patch_size)
    if spec.kind == "semseg":
        assert x.dim() == 3
        H, W = x.shape[1:]
        assert H % P == 0 and W % P == 0
        emb = p[prefix + ".class_emb.weight"][x]                            # :229  [B,H,W,E]
        tok = _patchify_linear(emb.permute(0, 3, 1, 2), p[prefix + ".proj.weight"], p[prefix + ".proj.bias"], P)
        pos = resized_posemb(p[prefix + ".pos_emb"], H // P, W // P, "bilinear")     # :235 (no align_corners arg)
    else:
        H, W = x.shape[2:]
        assert H % P == 0 and W % P == 0
        tok = _patchify_linear(x, p[prefix + ".proj.weight"], p[prefix + ".proj.bias"], P)   # :110
        pos = resized_posemb(p[prefix + ".pos_emb"], H // P, W // P, "bicubic")      # :113
    return tok + pos[None]


# --------------------------------------------------------------------------------------------------------------
# output adapter (multimae/output_adapters.py:160-282)
# --------------------------------------------------------------------------------------------------------------
def decode_task(enc, p, key: str, spec: DomainSpec, cfg: OracleConfig, token_counts: Dict[str, int],
                image_hw: Tuple[int, int], ids_keep, ids_restore):
    prefix = "output_adapters.%s" % key
    B = enc.shape[0]
    P = spec.patch(cfg.patch_size)
    H, W = image_hw
    nh, nw = H // (spec.stride_level * P), W // (spec.stride_level * P)
    G = cfg.num_global_tokens
    total = sum(token_counts.values())

    ctx = _lin(enc, p, prefix + ".proj_context")                            # :258
    body, glob = ctx[:, :ctx.shape[1] - G], ctx[:, ctx.shape[1] - G:]       # :190-193
    filler = p[prefix + ".mask_token"].expand(B, total - body.shape[1], -1)           # :196-198
    full = torch.cat([body, filler], dim=1)
    full = torch.gather(full, 1, ids_restore[:, :, None].expand(-1, -1, full.shape[2]))    # :201-202
    # :160-181 — every context task gets its task embedding + the bilinear-resized pos-emb (same grid for all)
    pos = resized_posemb(p[prefix + ".pos_emb"], nh, nw, "bilinear")
    embs = []
    start = {}
    off = 0
    for name, n in token_counts.items():
        start[name] = off
        off += n
        te = p.get(prefix + ".task_embeddings." + name)
        e = pos if te is None else pos + te.reshape(1, -1)
        assert e.shape[0] == n
        embs.append(e)
    full = full + torch.cat(embs, dim=0)[None]                              # :207
    task = spec.name
    if cfg.use_task_queries and task in token_counts:
        queries = full[:, start[task]:start[task] + token_counts[task]]     # :209-213
    else:                                                                   # :214-221 — mask-token queries
        queries = p[prefix + ".mask_token"].expand(B, nh * nw, -1) + pos[None]
        te = p.get(prefix + ".task_embeddings." + task)
        if te is not None:
            queries = queries + te.reshape(1, 1, -1)
    vis = torch.gather(full, 1, ids_keep[:, :, None].expand(-1, -1, full.shape[2]))   # :224-225
    context = torch.cat([vis, glob], dim=1)                                 # :229-230 (global token: no embedding)

    x = _cross_attention(_ln(queries, p, prefix + ".query_norm", cfg.eps),
                         _ln(context, p, prefix + ".context_norm", cfg.eps), p, prefix + ".decoder", cfg.dec_heads)  # :265
    x = x + _mlp(_ln(x, p, prefix + ".out_norm", cfg.eps), p, prefix + ".mlp")       # :266
    for i in range(cfg.dec_depth):                                          # :271
        x = _block(x, p, "%s.decoder_transformer.%d" % (prefix, i), cfg.dec_heads, cfg.eps)
    x = _lin(x, p, prefix + ".out_proj")                                    # :274
    C = spec.channels
    x = x.reshape(B, nh, nw, C, P, P).permute(0, 3, 1, 4, 2, 5).reshape(B, C, nh * P, nw * P)   # :277-280
    return x


def ids_from_task_masks(task_masks: List[torch.Tensor]):
    """Fixed-mask branch of MultiMAE.forward (multimae/multimae.py:334-338): visible tokens (mask 0) first.  The reference
    sorts with an unstable argsort and takes ONE visible count from the whole batch (right only when every sample keeps
    the same number, e.g. B = 1); here the sort is stable and the count is per sample, which yields the same predictions
    (the encoder and the decoders' context are permutation-invariant in the kept tokens)."""
    mask_all = torch.cat(list(task_masks), dim=1)
    ids_shuffle = torch.argsort(mask_all, dim=1, stable=True)
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
    n_vis = (mask_all == 0).sum(dim=1)
    assert bool((n_vis == n_vis[0]).all()), "every sample must keep the same number of visible tokens"
    return ids_shuffle[:, :int(n_vis[0])], ids_restore


# --------------------------------------------------------------------------------------------------------------
# whole forward (multimae/multimae.py:271-379)
# --------------------------------------------------------------------------------------------------------------
def forward(p: Dict[str, torch.Tensor], x: Dict[str, torch.Tensor], cfg: OracleConfig, ids_keep, ids_restore):
    """Returns (preds dict, encoder_tokens).  Masks are supplied (recorded or from sample_masks)."""
    first = cfg.in_domains[0]
    if "rgb" in x:
        H, W = x["rgb"].shape[2:]
    elif "semseg" in x:
        H, W = [s * SEMSEG.stride_level for s in x["semseg"].shape[1:]]
    else:
        H, W = x[first.name].shape[2:]
    tokens = {d.name: embed_domain(x[d.name], p, d, cfg) for d in cfg.in_domains if d.name in x}   # :312-316
    counts = {k: v.shape[1] for k, v in tokens.items()}
    seq = torch.cat(list(tokens.values()), dim=1)                           # :340
    B = seq.shape[0]
    seq = torch.gather(seq, 1, ids_keep[:, :, None].expand(-1, -1, seq.shape[2]))       # :343
    seq = torch.cat([seq, p["global_tokens"].expand(B, -1, -1)], dim=1)     # :346-347 (global token LAST)
    for i in range(cfg.depth):                                              # :350
        seq = _block(seq, p, "encoder.%d" % i, cfg.heads, cfg.eps)
    preds = {key: decode_task(seq, p, key, spec, cfg, counts, (H, W), ids_keep, ids_restore)
             for key, spec in cfg.out_tasks}                                # :357-377
    return preds, seq


# --------------------------------------------------------------------------------------------------------------
# masked losses (multimae/criterion.py:37-57, 84-114, 141-171)
# --------------------------------------------------------------------------------------------------------------
def _pixel_mask(mask, H, W, scale):
    B = mask.shape[0]
    nh, nw = H // scale, W // scale
    m = mask.reshape(B, nh, nw).float()
    return m.repeat_interleave(scale, 1).repeat_interleave(scale, 2)        # nearest upsample, integer factor


def _norm_pix_target(target, scale):                                        # :74-77, 89-95 — (p1 p2 c) order, unbiased var
    B, C, H, W = target.shape
    nh, nw = H // scale, W // scale
    t = target.reshape(B, C, nh, scale, nw, scale).permute(0, 2, 4, 3, 5, 1).reshape(B, nh * nw, scale * scale * C)
    t = (t - t.mean(-1, keepdim=True)) / torch.sqrt(t.var(-1, keepdim=True) + 1e-6)
    return t.reshape(B, nh, nw, scale, scale, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, H, W)


def _masked_mean(per_pixel, mask, scale):
    if mask is None:
        return per_pixel.mean()
    if int(mask.sum()) == 0:
        return torch.zeros((), dtype=per_pixel.dtype)
    H, W = per_pixel.shape[-2:]
    m = _pixel_mask(mask, H, W, scale)
    per_sample = (per_pixel * m).flatten(1).sum(1) / m.flatten(1).sum(1)    # 0/0 -> nan for empty samples
    return per_sample.nanmean()


def masked_mse(pred, target, mask=None, patch_size=16, stride=1, norm_pix=False):
    scale = patch_size // stride
    if norm_pix:
        target = _norm_pix_target(target, scale)
    return _masked_mean(((pred - target) ** 2).mean(1) if mask is not None else (pred - target) ** 2, mask, scale)


def masked_l1(pred, target, mask=None, patch_size=16, stride=1, norm_pix=False):
    scale = patch_size // stride
    if norm_pix:
        target = _norm_pix_target(target, scale)
    return _masked_mean((pred - target).abs().mean(1) if mask is not None else (pred - target).abs(), mask, scale)


def masked_ce(logits, target, mask=None, patch_size=16, stride=1, label_smoothing=0.0):
    scale = patch_size // stride
    logp = torch.log_softmax(logits, dim=1)
    nll = -logp.gather(1, target[:, None]).squeeze(1)
    if label_smoothing > 0:
        nll = (1 - label_smoothing) * nll + label_smoothing * (-logp.mean(1))
    return _masked_mean(nll, mask, scale)


def default_losses(cfg: OracleConfig):
    """Loss per output key as wired by main() (run_pretraining_multimae.py:321-330)."""
    out = {}
    for key, spec in cfg.out_tasks:
        if key == "norm_rgb":
            out[key] = lambda pr, tg, m, s=spec: masked_mse(pr, tg, m, cfg.patch_size, s.stride_level, norm_pix=True)
        elif spec.kind == "semseg":
            out[key] = lambda pr, tg, m, s=spec: masked_ce(pr, tg, m, cfg.patch_size, s.stride_level)
        elif spec.name == "depth":
            out[key] = lambda pr, tg, m, s=spec: masked_l1(pr, tg, m, cfg.patch_size, s.stride_level)
        else:
            out[key] = lambda pr, tg, m, s=spec: masked_mse(pr, tg, m, cfg.patch_size, s.stride_level)
    return out


def step_losses(p, x, cfg: OracleConfig, task_masks: Dict[str, torch.Tensor], ids_keep, ids_restore, targets=None):
    """One train_one_epoch body up to the loss (run_pretraining_multimae.py:494-523): returns (losses, preds).
    `x` is the input_dict handed to the model (:494-498); `targets` is tasks_dict (defaults to x) — it may hold tasks that
    are reconstructed without being fed, whose loss then runs without a mask (masks.get(task, None), :520)."""
    targets = x if targets is None else targets
    preds, _ = forward(p, x, cfg, ids_keep, ids_restore)
    fns = default_losses(cfg)
    losses = {}
    for key, spec in cfg.out_tasks:
        mask = task_masks.get(spec.name)
        losses[key] = fns[key](preds[key].float(), targets[spec.name], mask)
    return losses, preds


def standardize_depth(depth: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Truncated depth standardisation, run_pretraining_multimae.py:487-492: per sample, sort the flattened map, drop the
    bottom and top 10 % of the values, standardise the whole map with the mean / unbiased variance of the rest."""
    flat = depth.reshape(depth.shape[0], -1)
    trunc = torch.sort(flat, dim=1)[0]                                       # :489
    trunc = trunc[:, int(0.1 * trunc.shape[1]): int(0.9 * trunc.shape[1])]   # :490
    shape = (-1,) + (1,) * (depth.dim() - 1)
    return (depth - trunc.mean(dim=1).reshape(shape)) / torch.sqrt(trunc.var(dim=1).reshape(shape) + eps)   # :491


def grad_norm(grads) -> torch.Tensor:
    """utils/native_scaler.py:49-62 — L2 norm of per-tensor L2 norms."""
    return torch.norm(torch.stack([g.detach().norm(2) for g in grads if g is not None]), 2)


# --------------------------------------------------------------------------------------------------------------
# parameter construction (shapes of SURVEY.md §A.1; init follows multimae/multimae.py:89-125 in spirit — exact RNG
# parity with the reference's init is not needed: tests copy the reference's state_dict)
# --------------------------------------------------------------------------------------------------------------
def init_params(cfg: OracleConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}

    def xavier(out_f, in_f, fan_out=None):
        bound = math.sqrt(6.0 / ((fan_out or out_f) + in_f))
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound

    def linear(prefix, out_f, in_f, fan_out=None):
        p[prefix + ".weight"] = xavier(out_f, in_f, fan_out)
        p[prefix + ".bias"] = torch.zeros(out_f)

    def norm(prefix, d):
        p[prefix + ".weight"] = torch.ones(d)
        p[prefix + ".bias"] = torch.zeros(d)

    def block(prefix, d):
        norm(prefix + ".norm1", d)
        linear(prefix + ".attn.qkv", 3 * d, d, fan_out=d)
        linear(prefix + ".attn.proj", d, d)
        norm(prefix + ".norm2", d)
        linear(prefix + ".mlp.fc1", 4 * d, d)
        linear(prefix + ".mlp.fc2", d, 4 * d)

    D, Dd, G = cfg.dim, cfg.dec_dim, cfg.posemb_grid
    p["global_tokens"] = torch.randn(1, cfg.num_global_tokens, D, generator=g).clamp(-2, 2) * 0.02
    for d in cfg.in_domains:
        pre = "input_adapters." + d.name
        P = d.patch(cfg.patch_size)
        p[pre + ".pos_emb"] = sincos_posemb(G, G, D)
        cin = d.dim_class_emb if d.kind == "semseg" else d.channels
        if d.kind == "semseg":
            p[pre + ".class_emb.weight"] = torch.randn(d.channels, d.dim_class_emb, generator=g).clamp(-2, 2) * 0.02
        p[pre + ".proj.weight"] = xavier(D, cin * P * P).reshape(D, cin, P, P)
        p[pre + ".proj.bias"] = torch.zeros(D)
    for i in range(cfg.depth):
        block("encoder.%d" % i, D)
    for key, spec in cfg.out_tasks:
        pre = "output_adapters." + key
        P = spec.patch(cfg.patch_size)
        p[pre + ".mask_token"] = torch.zeros(1, 1, Dd)
        p[pre + ".pos_emb"] = sincos_posemb(G, G, Dd)
        for name in (cfg.context_tasks if cfg.context_tasks is not None else [d.name for d in cfg.in_domains]):
            p[pre + ".task_embeddings." + name] = torch.randn(1, 1, Dd, generator=g).clamp(-2, 2) * 0.02
        linear(pre + ".proj_context", Dd, D)
        for n in ("context_norm", "query_norm", "out_norm"):
            norm(pre + "." + n, Dd)
        linear(pre + ".decoder.q", Dd, Dd)
        linear(pre + ".decoder.kv", 2 * Dd, Dd, fan_out=Dd)
        linear(pre + ".decoder.proj", Dd, Dd)
        linear(pre + ".mlp.fc1", 4 * Dd, Dd)
        linear(pre + ".mlp.fc2", Dd, 4 * Dd)
        for i in range(cfg.dec_depth):
            block("%s.decoder_transformer.%d" % (pre, i), Dd)
        linear(pre + ".out_proj", spec.channels * P * P, Dd)
    return p


FROZEN_SUFFIXES = (".pos_emb",)


def trainable(p: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k: v for k, v in p.items() if not k.endswith(FROZEN_SUFFIXES)}


# --------------------------------------------------------------------------------------------------------------
# synthetic inputs: one fixed generator, CPU, moved to device by the caller
# --------------------------------------------------------------------------------------------------------------
def synthetic_inputs(cfg: OracleConfig, batch: int, image_size: int = 224, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    x = {}
    for d in cfg.in_domains:
        s = image_size // d.stride_level
        if d.kind == "semseg":
            x[d.name] = torch.randint(0, d.channels, (batch, s, s), generator=g)
        else:
            x[d.name] = torch.randn(batch, d.channels, s, s, generator=g)
    return x


def synthetic_mask_draws(cfg: OracleConfig, batch: int, image_size: int = 224, seed: int = 1, alphas: float = 1.0):
    """(shares, noises, noise_all) as generate_random_masks would draw them (multimae/multimae.py:182-204)."""
    g = torch.Generator().manual_seed(seed)
    T = len(cfg.in_domains)
    gamma = torch.distributions.Gamma(torch.full((T,), float(alphas)), torch.ones(T))
    torch.manual_seed(seed)
    draw = gamma.sample((batch,))
    shares = draw / draw.sum(-1, keepdim=True)                               # Dirichlet(alpha) draw
    n_tok = [(image_size // (d.stride_level * d.patch(cfg.patch_size))) ** 2 for d in cfg.in_domains]
    noises = [torch.rand(batch, n, generator=g) for n in n_tok]
    noise_all = torch.rand(batch, sum(n_tok), generator=g)
    return shares, noises, noise_all
