"""Flat-buffer AdamW: parameters live as views of one fp32 buffer, gradients in the model's
GradArena, so the update of all ~98 M parameters is ONE kernel (mmae_adamw_step) instead of per-tensor loops, and a
non-finite step is skipped on the device.  Same update rule / hyper-parameter semantics as torch.optim.AdamW as the
reference builds it (utils/optim_factory.py:155-174; betas (0.9, 0.95), weight_decay 0.05 on every parameter)."""
import torch

from . import functional as Fn


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05):
        params = [p for p in model.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, lr_scale=1.0))
        arena = model.own_gradients(True)
        self.mmae_arena = arena
        dev = arena.flat.device
        # re-home the parameters into one flat buffer laid out exactly like the gradient arena
        self.flat_params = torch.zeros_like(arena.flat)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if not p.requires_grad:
                    continue
                o, numel, shape = arena.offsets[n]
                view = self.flat_params[o:o + numel].view(shape)
                view.copy_(p.data.to(dev))
                p.data = view
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self._step = 0

    @torch.no_grad()
    def fused_step(self, found_inf=None):
        g = self.param_groups[0]
        self._step += 1
        Fn.adamw_step(self.flat_params, self.mmae_arena.flat, self.exp_avg, self.exp_avg_sq, g["lr"] * g.get("lr_scale", 1.0),
                      g["betas"], g["eps"], g["weight_decay"], self._step, found_inf)

    @torch.no_grad()
    def step(self, closure=None):
        self.fused_step(None)

    def zero_grad(self, set_to_none=True):
        """Gradients are zeroed by the model at the start of every forward (one memset of the arena)."""
        return None
