"""SpatialOutputAdapter — the pre-training decoder — with the reference's constructor / state_dict / forward contract
(multimae/output_adapters.py:33-282), executed as DecoderHeadFunction -> Block x depth -> DecoderTailFunction.

The fine-tuning heads of the reference (Linear / Segmenter / ConvNeXt / DPT adapters) are outside the pre-training hot
path and are not provided."""
from functools import partial
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import functional as Fn
from .input_adapters import _PosEmbCache
from .multimae_utils import Block, CrossAttention, Mlp, build_2d_sincos_posemb, pair, trunc_normal_


class SpatialOutputAdapter(nn.Module, _PosEmbCache):
    """Cross-attention adapter for spatial outputs, like images or feature maps."""

    def __init__(self, num_channels: int, stride_level: int, patch_size_full: Union[int, Tuple[int, int]],
                 dim_tokens_enc: Optional[int] = None, dim_tokens: int = 256, depth: int = 0,
                 learnable_pos_emb: int = False, image_size: Union[int, Tuple[int]] = 224, mlp_ratio: int = 4.0,
                 num_heads: int = 8, qkv_bias: bool = True, drop_rate: float = 0.0, attn_drop_rate: float = 0.0,
                 drop_path_rate: float = 0.0, norm_layer: nn.Module = partial(nn.LayerNorm, eps=1e-6),
                 use_task_queries: bool = True, task: Optional[str] = None, context_tasks: Optional[list] = None,
                 use_xattn: bool = True):
        super().__init__()
        self.num_channels = num_channels
        self.stride_level = stride_level
        self.patch_size_full = pair(patch_size_full)
        self.dim_tokens_enc = dim_tokens_enc
        self.dim_tokens = dim_tokens
        self.learnable_pos_emb = learnable_pos_emb
        self.image_size = pair(image_size)
        self.use_task_queries = use_task_queries
        self.task = task
        self.use_xattn = use_xattn
        self.num_heads = num_heads
        self.P_H = max(1, self.patch_size_full[0] // stride_level)
        self.P_W = max(1, self.patch_size_full[1] // stride_level)
        assert self.P_H == self.P_W, "multimae_b200: square patches only"

        self.task_embeddings = None
        if context_tasks is not None:
            self.task_embeddings = nn.ParameterDict(
                {t: nn.Parameter(torch.zeros(1, 1, self.dim_tokens)) for t in context_tasks})
            for emb in self.task_embeddings.values():
                trunc_normal_(emb, std=0.02)

        self.mask_token = nn.Parameter(torch.zeros(1, 1, self.dim_tokens))

        h_posemb = self.image_size[0] // (self.stride_level * self.P_H)
        w_posemb = self.image_size[1] // (self.stride_level * self.P_W)
        if not self.learnable_pos_emb:
            self.pos_emb = nn.Parameter(build_2d_sincos_posemb(h=h_posemb, w=w_posemb, embed_dim=self.dim_tokens),
                                        requires_grad=False)
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, h_posemb, w_posemb, self.dim_tokens))
            trunc_normal_(self.pos_emb, std=0.02)

        if self.use_xattn:
            self.decoder = CrossAttention(dim=self.dim_tokens, num_heads=num_heads, qkv_bias=qkv_bias,
                                          attn_drop=attn_drop_rate, proj_drop=drop_rate)
            self.context_norm = norm_layer(self.dim_tokens)
            self.query_norm = norm_layer(self.dim_tokens)
            self.out_norm = norm_layer(self.dim_tokens)
            self.mlp_hidden = int(self.dim_tokens * mlp_ratio)
            self.mlp = Mlp(in_features=self.dim_tokens, hidden_features=self.mlp_hidden)

        if depth > 0:
            dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
            self.decoder_transformer = nn.Sequential(*[
                Block(dim=self.dim_tokens, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                      attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer) for i in range(depth)])
        else:
            self.decoder_transformer = nn.Identity()

        self.dim_patch = self.num_channels * self.P_H * self.P_W
        self.out_proj = nn.Linear(self.dim_tokens, self.dim_patch)
        self._bound = None
        if self.dim_tokens_enc is not None:
            self.init(dim_tokens_enc=dim_tokens_enc)

    def init(self, dim_tokens_enc: int = 768):
        self.dim_tokens_enc = dim_tokens_enc
        self.proj_context = nn.Linear(self.dim_tokens_enc, self.dim_tokens)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_emb", "mask_token", "task_embeddings"}

    # ------------------------------------------------------------------------------------------------------------
    def bind(self, arena, prefix, on_grads_ready=None):
        self._bound = dict(arena=arena, prefix=prefix, on_grads_ready=on_grads_ready)
        if isinstance(self.decoder_transformer, nn.Sequential):
            for i, blk in enumerate(self.decoder_transformer):
                blk.bind(arena, "%sdecoder_transformer.%d." % (prefix, i), on_grads_ready)

    def _head_params(self):
        return (self.proj_context.weight, self.proj_context.bias, self.mask_token, self.context_norm.weight,
                self.context_norm.bias, self.query_norm.weight, self.query_norm.bias, self.out_norm.weight,
                self.out_norm.bias, self.decoder.q.weight, self.decoder.q.bias, self.decoder.kv.weight,
                self.decoder.kv.bias, self.decoder.proj.weight, self.decoder.proj.bias, self.mlp.fc1.weight,
                self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias)

    def forward(self, encoder_tokens: torch.Tensor, input_info: Dict, ids_keep: torch.Tensor, ids_restore: torch.Tensor,
                fp32: bool = False, shared_ctx: Dict = None):
        """`fp32` (beyond the reference signature): run the whole adapter in the fp32 tier - what the reference does for the
        adapters listed in `fp32_output_adapters` by calling them outside autocast (multimae/multimae.py:367-377).
        `shared_ctx` (set by MultiMAE._decode): proj_context was already applied for all adapters in one GEMM
        (functional.SharedContextFunction); dict(ctx=[B*Nc, sum Dd] tensor, offset, ld, state, enc_shape)."""
        assert self.dim_tokens_enc is not None, "Need to call init(dim_tokens_enc) function first"
        if not self.use_xattn:
            raise NotImplementedError("multimae_b200: use_xattn=False is outside the pre-training hot path")
        if self.learnable_pos_emb:
            raise NotImplementedError("multimae_b200: learnable_pos_emb=True is outside the pre-training hot path")
        H, W = input_info["image_size"]
        nh = H // (self.stride_level * self.P_H)
        nw = W // (self.stride_level * self.P_W)
        tasks = list(input_info["tasks"].keys())
        task_names = list(tasks)
        task_embs = [self.task_embeddings[t] if (self.task_embeddings is not None and t in self.task_embeddings) else None
                     for t in tasks]
        if self.use_task_queries and self.task in tasks:
            # queries = this task's rows of the restored, embedded context (multimae/output_adapters.py:209-213)
            query_mode, own_task = 0, tasks.index(self.task)
        else:
            # queries = mask_token + pos_emb (+ this task's embedding when there is one) (:214-221): use_task_queries=False,
            # or a task that is reconstructed without being an input (e.g. --in_domains rgb --out_domains rgb-depth-semseg)
            query_mode, own_task = 1, -1
            if self.task_embeddings is not None and self.task in self.task_embeddings:
                if self.task in tasks:
                    own_task = tasks.index(self.task)
                else:                                        # an embedding that belongs to none of the given inputs
                    if len(tasks) >= Fn.L.MAX_TASKS:
                        raise Fn.L.MmaeError("multimae_b200: no free task slot for the query task embedding")
                    own_task = len(tasks)
                    task_names.append(self.task)
                    task_embs.append(self.task_embeddings[self.task])
        tok_offset = [input_info["tasks"][t]["start_idx"] for t in tasks] + [input_info["num_task_tokens"]]
        for t in tasks:
            assert input_info["tasks"][t]["num_tokens"] == nh * nw, "context tasks must share the adapter's patch grid"
        if self._bound is None or self._bound["arena"].flat.device != encoder_tokens.device:
            self.bind(Fn.GradArena([(n, p) for n, p in self.named_parameters() if p.requires_grad],
                                   encoder_tokens.device), "")
            self._own_arena = True
        if getattr(self, "_own_arena", False) and torch.is_grad_enabled():
            self._bound["arena"].zero_()
        head_meta = dict(self._bound, dim=self.dim_tokens, num_global=input_info.get("num_global_tokens", 0),
                         num_queries=nh * nw, tok_offset=tok_offset, own_task=own_task, query_mode=query_mode,
                         heads=self.num_heads, hidden=self.mlp_hidden, eps=self.query_norm.eps,
                         pos=self._resized_pos(nh, nw, "bilinear"), task_names=task_names, fp32=bool(fp32))
        head_params = self._head_params()
        head_in = encoder_tokens
        if shared_ctx is not None:
            assert not fp32, "the shared context projection is the half-precision tier"
            head_meta["shared"] = {k: shared_ctx[k] for k in ("offset", "ld", "state", "enc_shape")}
            head_in = shared_ctx["ctx"]
            head_params = (None,) + head_params[1:]          # proj_context.weight: used and differentiated by the shared GEMM
        x = Fn.DecoderHeadFunction.apply(head_in, head_meta, ids_keep, ids_restore, *head_params, *task_embs)
        if fp32 and isinstance(self.decoder_transformer, nn.Sequential):
            for blk in self.decoder_transformer:
                x = blk(x, fp32=True)
        elif isinstance(self.decoder_transformer, nn.Sequential):
            x = Fn.block_stack(self.decoder_transformer, x)
        else:
            x = self.decoder_transformer(x)
        tail_meta = dict(self._bound, nh=nh, nw=nw, channels=self.num_channels, patch=self.P_H, fp32=bool(fp32))
        return Fn.DecoderTailFunction.apply(x, tail_meta, self.out_proj.weight, self.out_proj.bias)
