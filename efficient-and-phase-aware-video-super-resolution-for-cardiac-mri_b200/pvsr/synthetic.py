"""Synthetic ACDCSR / DSB15SR-shaped cine sequences (the datasets are not available offline).

Shapes and conventions follow the reference data path:
  - frames normalised like transforms.Normalize (configs/test/refine_net/exp1_x4.yaml:12-15), here plain randn;
  - the cardiac cycle is circularly padded by U frames on each side exactly like
    AcdcVSRRefineNetDataset.__getitem__ (src/data/datasets/acdc_vsr_refinenet_dataset.py:74-87);
  - positional code = two half cosines split at end-systole (src/gen_positional_encoding.py:35-38).
"""
import numpy as np
import torch

ACDC_X4 = dict(h=54, w=63, scale=4)
ACDC_X3 = dict(h=72, w=84, scale=3)
ACDC_X2 = dict(h=108, w=126, scale=2)
DSB15_X4 = dict(h=63, w=48, scale=4)


def positional_code(T, end_systole):
    y1 = np.cos(np.linspace(0, np.pi, end_systole, endpoint=False))
    y2 = np.cos(np.linspace(np.pi, np.pi * 2, T - end_systole, endpoint=False))
    return np.concatenate((y1, y2)).astype(np.float32)


def circular_window(frames, T, U):
    """Test-mode slice [T-U, 2T+U) of the cycle tiled three times."""
    tiled = list(frames) * 3
    return tiled[T - U:2 * T + U]


def cine_batch(batch, T=30, U=6, h=54, w=63, scale=4, seed=1234, end_systole=11, with_targets=False):
    """Returns (lr_frames: list of T+2U tensors (batch,1,h,w), pos_codes (batch,T+2U,1)[, hr targets list of T])."""
    g = torch.Generator().manual_seed(seed)
    frames = [torch.randn(batch, 1, h, w, generator=g) for _ in range(T)]
    inputs = circular_window(frames, T, U)
    code = torch.from_numpy(positional_code(T, end_systole))
    pos = torch.stack(circular_window(list(code), T, U)).view(1, -1, 1).repeat(batch, 1, 1).contiguous()
    if with_targets:
        hr = [torch.randn(batch, 1, h * scale, w * scale, generator=g) for _ in range(T)]
        return inputs, pos, hr
    return inputs, pos
