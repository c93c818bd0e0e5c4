"""Flat-buffer AdamW: parameters live as views of one fp32 buffer, gradients in the model's
GradArena, so the update of all ~98 M parameters is ONE kernel (mmae_adamw_step) instead of per-tensor loops, and a
non-finite step is skipped on the device.  Same update rule / hyper-parameter semantics as torch.optim.AdamW as the
reference builds it (utils/optim_factory.py:155-174; betas (0.9, 0.95), weight_decay 0.05 on every parameter).

Checkpoints: `state_dict()` / `load_state_dict()` use torch.optim.AdamW's per-parameter layout ({step, exp_avg,
exp_avg_sq} per parameter index, sliced out of / copied into the flat moment buffers), so utils.save_model /
auto_load_model (utils/checkpoint.py:85,124-133) resume this optimizer - and a checkpoint written by the stock AdamW over
the same parameter order loads here, and vice versa."""
import torch

from . import _lib as L
from . import functional as Fn


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        params = [p for _, p in named]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, lr_scale=1.0))
        arena = model.own_gradients(True)
        self.mmae_arena = arena
        dev = arena.flat.device
        # re-home the parameters into one flat buffer laid out exactly like the gradient arena
        self.flat_params = torch.zeros_like(arena.flat)
        with torch.no_grad():
            for n, p in named:
                o, numel, shape = arena.offsets[n]
                view = self.flat_params[o:o + numel].view(shape)
                view.copy_(p.data.to(dev))
                p.data = view
        # bf16 twin of the parameters: the GEMM weight operands are read from it (no per-step casts of ~117 tensors); the
        # update kernel keeps it current, ensure_mirror_fresh() catches every other modification through torch
        self.flat_bf16 = None
        self._mirror_version = None
        self._invalidations = 0
        self._flat_views = params
        self._named = named
        if dev.type == "cuda":
            self.flat_bf16 = torch.empty(self.flat_params.numel(), dtype=torch.bfloat16, device=dev)
            L.check(L.lib().mmae_weight_mirror_register(self.flat_params.data_ptr(), self.flat_bf16.data_ptr(),
                                                        self.flat_params.numel()), "mmae_weight_mirror_register")
            Fn.register_mirror(self)
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        # {learning rate, step count} on the device: the update kernel reads them there, so a captured CUDA graph replays
        # with the current schedule value and an advancing bias correction
        self._dyn = torch.zeros(2, dtype=torch.float32, device=dev)
        self._lr_host = torch.zeros(1, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(1)
        self._lr_synced = None
        self._bind_state()
        self.sync_hyperparams()

    # ------------------------------------------------------------------------------------------------------------
    # optimizer state in torch.optim.AdamW's layout, as views of the flat buffers (checkpoint / resume contract)
    # ------------------------------------------------------------------------------------------------------------
    def _bind_state(self):
        arena = self.mmae_arena
        step = self._dyn[1]                      # 0-dim view of the device-side step counter, shared by every entry
        for n, p in self._named:
            o, numel, shape = arena.offsets[n]
            self.state[p] = {"step": step, "exp_avg": self.exp_avg[o:o + numel].view(shape),
                             "exp_avg_sq": self.exp_avg_sq[o:o + numel].view(shape)}

    def load_state_dict(self, state_dict):
        """torch.optim.AdamW-layout state -> flat moment buffers + device step counter; hyper-parameters as usual."""
        super().load_state_dict(state_dict)      # validates sizes, restores param_groups, materialises per-parameter state
        loaded_step = None
        with torch.no_grad():
            for n, p in self._named:
                st = self.state.get(p)
                if not st:
                    continue
                o, numel, shape = self.mmae_arena.offsets[n]
                if "exp_avg" in st:
                    self.exp_avg[o:o + numel].view(shape).copy_(st["exp_avg"])
                if "exp_avg_sq" in st:
                    self.exp_avg_sq[o:o + numel].view(shape).copy_(st["exp_avg_sq"])
                if "step" in st and loaded_step is None:
                    loaded_step = float(st["step"])
            if loaded_step is not None:
                self._dyn[1].fill_(loaded_step)
        self._bind_state()
        self._lr_synced = None
        self.sync_hyperparams()
        self.invalidate_mirror()

    def _params_version(self):
        # in-place torch ops on a Parameter (copy_, load_state_dict, init) bump its counter; our kernels do not
        return self._invalidations + sum(p._version for p in self._flat_views)

    def invalidate_mirror(self):
        """Call after modifying parameters in a way autograd's version counters do not see (writes through `.data`)."""
        self._invalidations += 1

    def ensure_mirror_fresh(self):
        """Re-cast the whole flat buffer (one kernel) if the parameters were modified through torch since the twin was
        last written.  The update kernel itself writes the twin and does not touch the version counters.  Called by the
        model at the start of every forward."""
        if self.flat_bf16 is None:
            return
        ver = self._params_version()
        if self._mirror_version == ver:
            return
        L.check(L.lib().mmae_cast_f32_to_bf16(self.flat_params.data_ptr(), self.flat_bf16.data_ptr(),
                                              self.flat_params.numel(), L.current_stream()), "mmae_cast_f32_to_bf16")
        self._mirror_version = ver

    def release_mirror(self):
        if getattr(self, "flat_bf16", None) is not None:
            try:
                L.lib().mmae_weight_mirror_register(self.flat_params.data_ptr(), None, 0)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass
            self.flat_bf16 = None

    def __del__(self):
        self.release_mirror()

    def sync_hyperparams(self):
        """Push the current param_group learning rate to the device scalar (call before a graph replay)."""
        g = self.param_groups[0]
        lr = g["lr"] * g.get("lr_scale", 1.0)
        if lr == self._lr_synced:
            return                                  # nothing to copy: keeps a replayed step free of H2D traffic
        self._lr_host[0] = lr
        self._dyn[0:1].copy_(self._lr_host, non_blocking=True)
        self._lr_synced = lr

    @torch.no_grad()
    def fused_step(self, found_inf=None):
        """One kernel over the flat buffers.  The step counter lives on the device and is advanced by the call itself,
        only when the step is taken (found_inf == 0): a skipped step leaves moments, parameters AND the bias correction
        untouched, like GradScaler.step."""
        g = self.param_groups[0]
        if not (self._dyn.is_cuda and torch.cuda.is_current_stream_capturing()):
            self.sync_hyperparams()
        Fn.adamw_step(self.flat_params, self.mmae_arena.flat, self.exp_avg, self.exp_avg_sq, g["lr"] * g.get("lr_scale", 1.0),
                      g["betas"], g["eps"], g["weight_decay"], 1, found_inf, dyn=self._dyn)

    @torch.no_grad()
    def step(self, closure=None):
        self.fused_step(None)

    def zero_grad(self, set_to_none=True):
        """Gradients alias the model's arena, which the model zeroes (one memset) at the start of the next training
        forward - unless an accumulation is in progress (NativeScalerWithGradNormCount(update_grad=False)); after a step the
        accumulation is over."""
        self.mmae_arena.accumulating = False
        return None
