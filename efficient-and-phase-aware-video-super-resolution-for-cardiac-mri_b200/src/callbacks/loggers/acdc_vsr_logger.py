"""HR | SR panels of the last frame of the last batch (reference src/callbacks/loggers/acdc_vsr_logger.py:7-30)."""
import torch

from .base_logger import BaseLogger


def _column(frames):
    """(N, 1, H, W) -> one (1, N*H + pad, W) column, every sample min-max scaled (what torchvision's
    make_grid(nrow=1, normalize=True, scale_each=True, pad_value=1) shows, without needing torchvision)."""
    rows = []
    for f in frames.detach().float():
        lo, hi = f.min(), f.max()
        rows.append((f - lo) / (hi - lo).clamp_min(1e-5))
        rows.append(torch.ones(1, 2, f.shape[-1], device=f.device))
    return torch.cat(rows[:-1], dim=1)


class AcdcVSRLogger(BaseLogger):
    def _add_images(self, epoch, train_batch, train_outputs, valid_batch, valid_outputs):
        for tag, batch, outputs in (('train', train_batch, train_outputs), ('valid', valid_batch, valid_outputs)):
            panel = torch.cat([_column(batch['hr_imgs'][-1]), _column(outputs[-1])], dim=-1)
            self._add_image(tag, panel, epoch)
