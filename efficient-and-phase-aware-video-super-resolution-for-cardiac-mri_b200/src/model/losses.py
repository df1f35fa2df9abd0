"""Custom losses resolvable by name from the `losses:` config section (reference src/main.py:63-70 falls back to
src.model.losses when torch.nn has no such loss).  The RefineNet configs only use torch.nn.L1Loss
(configs/train/refine_net/exp1_x4.yaml:48-50); the extra entries exist so foreign configs fail with a clear message
instead of an AttributeError."""
import torch
import torch.nn as nn

__all__ = ['CharbonnierLoss', 'HuberLoss']


class CharbonnierLoss(nn.Module):
    """sqrt((x - y)^2 + eps^2) averaged over all elements (a smooth L1; reference src/model/losses.py:23-34)."""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps2 = eps * eps

    def forward(self, output, target):
        d = output - target
        return torch.sqrt(d * d + self.eps2).mean()


class HuberLoss(nn.Module):
    """Quadratic below `delta`, linear above (reference src/model/losses.py:5-20)."""

    def __init__(self, delta=1.0):
        super().__init__()
        self.delta = delta

    def forward(self, output, target):
        a = (output - target).abs()
        quad = torch.clamp(a, max=self.delta)
        return (0.5 * quad * quad + self.delta * (a - quad)).mean()
