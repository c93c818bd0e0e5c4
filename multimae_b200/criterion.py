"""Masked reconstruction losses with the reference's constructor / call contract (multimae/criterion.py), each executed
as one fused forward kernel + one fused backward kernel with no host synchronisation."""
import torch.nn as nn

from . import functional as Fn


class MaskedCrossEntropyLoss(nn.Module):
    """Cross-entropy loss with masking (multimae/criterion.py:23-57)."""

    def __init__(self, patch_size: int = 16, stride: int = 1, label_smoothing: float = 0.0):
        super().__init__()
        self.patch_size = patch_size
        self.stride = stride
        self.scale_factor = patch_size // stride
        self.label_smoothing = label_smoothing

    def forward(self, input, target, mask=None):
        return Fn.MaskedLossFunction.apply(input, target, mask, 2, False, self.scale_factor, self.label_smoothing)


class MaskedMSELoss(nn.Module):
    """MSE loss with masking and optional per-patch target normalisation (multimae/criterion.py:60-114)."""

    def __init__(self, patch_size: int = 16, stride: int = 1, norm_pix=False):
        super().__init__()
        self.patch_size = patch_size
        self.stride = stride
        self.scale_factor = patch_size // stride
        self.norm_pix = norm_pix

    def forward(self, input, target, mask=None):
        return Fn.MaskedLossFunction.apply(input, target, mask, 0, bool(self.norm_pix), self.scale_factor, 0.0)


class MaskedL1Loss(nn.Module):
    """L1 loss with masking and optional per-patch target normalisation (multimae/criterion.py:117-171)."""

    def __init__(self, patch_size: int = 16, stride: int = 1, norm_pix=False):
        super().__init__()
        self.patch_size = patch_size
        self.stride = stride
        self.scale_factor = patch_size // stride
        self.norm_pix = norm_pix

    def forward(self, input, target, mask=None):
        return Fn.MaskedLossFunction.apply(input, target, mask, 1, bool(self.norm_pix), self.scale_factor, 0.0)
