"""DRFNet - drop-in for the reference's src/model/nets/drf_net.py (class at :8, forward at :38-49).

Same constructor, same `forward(inputs)` contract (a list of T tensors (N, 1, h, w) -> a list of T tensors
(N, 1, s*h, s*w)), same `state_dict` keys / shapes (`in_block.conv{1,2}` / `prelu{1,2}`, `f_block.in_block.*`,
`f_block.up_blocks.{i}.*`, `f_block.down_blocks.{i}.*`, `f_block.out_block.*`, `out_block.conv{k}`) and the same
construction order, so `torch.manual_seed(s)` + construction yields the reference's weights and reference checkpoints
load with strict=True.  The sub-modules only hold the fp32 master parameters; the arithmetic runs on the RefineNet
conv core (tcgen05 implicit-GEMM kernels behind include/pvsr.h) through pvsr.drf_engine.DRFEngine - including the
projection units' ConvTranspose2d / strided Conv2d, which become 3x3 convs over phase-stacked LR-grid tensors.
No CPU / PyTorch fallback.
"""
import torch
import torch.nn as nn

from .base_net import BaseNet


def _prelu():
    return nn.PReLU(num_parameters=1, init=0.2)


class _ParamFBlock(nn.Module):
    """Parameter holder of _FBlock (drf_net.py:61-116)."""

    def __init__(self, num_features, num_groups, kernel_size, stride, padding):
        super().__init__()
        F = num_features
        geo = dict(kernel_size=kernel_size, stride=stride, padding=padding)
        self.in_block = nn.Sequential()
        self.in_block.add_module('conv', nn.Conv2d(F * 2, F, kernel_size=1))
        self.in_block.add_module('prelu', _prelu())
        self.up_blocks = nn.ModuleList()
        self.down_blocks = nn.ModuleList()
        for i in range(num_groups):
            up, down = nn.Sequential(), nn.Sequential()
            if i == 0:
                up.add_module('deconv', nn.ConvTranspose2d(F, F, **geo))
                up.add_module('prelu', _prelu())
                self.up_blocks.append(up)
                down.add_module('conv', nn.Conv2d(F, F, **geo))
                down.add_module('prelu', _prelu())
                self.down_blocks.append(down)
            else:
                up.add_module('conv1', nn.Conv2d(F * (i + 1), F, kernel_size=1))
                up.add_module('prelu1', _prelu())
                up.add_module('deconv2', nn.ConvTranspose2d(F, F, **geo))
                up.add_module('prelu2', _prelu())
                self.up_blocks.append(up)
                down.add_module('conv1', nn.Conv2d(F * (i + 1), F, kernel_size=1))
                down.add_module('prelu1', _prelu())
                down.add_module('conv2', nn.Conv2d(F, F, **geo))
                down.add_module('prelu2', _prelu())
                self.down_blocks.append(down)
        self.out_block = nn.Sequential()
        self.out_block.add_module('conv', nn.Conv2d(F * num_groups, F, kernel_size=1))
        self.out_block.add_module('prelu', _prelu())


class DRFNet(BaseNet):
    """
    Args:
        in_channels (int): The input channels (1 on this path).
        out_channels (int): The output channels (1 on this path).
        num_features (int): The number of the internal feature maps (a multiple of 64, at most 256).
        num_groups (int): The number of the projection groups in the feedback block (at most 10).
        upscale_factor (int): The upscale factor (2, 3, 4 or 8).
    """

    def __init__(self, in_channels, out_channels, num_features, num_groups, upscale_factor):
        super().__init__()
        if upscale_factor not in [2, 3, 4, 8]:
            raise ValueError(f'The upscale factor should be 2, 3, 4 or 8. Got {upscale_factor}.')
        if in_channels != 1 or out_channels != 1:
            raise ValueError('The B200 path implements the single-channel cine-MRI configuration '
                             f'(in_channels = out_channels = 1). Got {in_channels}, {out_channels}.')
        if num_features % 64 != 0 or not 64 <= num_features <= 256:
            raise ValueError(f'The B200 path needs num_features in (64, 128, 192, 256). Got {num_features}.')
        if not 1 <= num_groups <= 10:
            raise ValueError(f'The B200 path reads at most 10 concatenated maps per launch (num_groups <= 10). Got {num_groups}.')
        from pvsr.drf_engine import PROJECTION, up_factors
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_features, self.num_groups, self.upscale_factor = num_features, num_groups, upscale_factor
        F = num_features

        self.in_block = nn.Sequential()
        self.in_block.add_module('conv1', nn.Conv2d(in_channels, 4 * F, kernel_size=3, padding=1))
        self.in_block.add_module('prelu1', _prelu())
        self.in_block.add_module('conv2', nn.Conv2d(4 * F, F, kernel_size=1))
        self.in_block.add_module('prelu2', _prelu())
        self.f_block = _ParamFBlock(F, num_groups, *PROJECTION[upscale_factor])
        self.out_block = nn.Sequential()
        factors = up_factors(upscale_factor)
        for i, r in enumerate(factors):
            self.out_block.add_module(f'conv{i + 1}', nn.Conv2d(F, r * r * F, kernel_size=3, padding=1))
            self.out_block.add_module(f'pixelshuffle{i + 1}', nn.PixelShuffle(r))
        self.out_block.add_module(f'conv{len(factors) + 1}', nn.Conv2d(F, out_channels, kernel_size=3, padding=1))
        self.reuse_output_buffers = False
        self._engine = None

    @property
    def engine(self):
        if self._engine is None:
            from pvsr.drf_engine import DRFEngine
            self._engine = DRFEngine(self)
        return self._engine

    def forward(self, inputs):
        inputs = list(inputs)
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            from pvsr.drf_engine import drf_train_forward
            return drf_train_forward(self, inputs)
        outs, _ = self.engine.forward(inputs, train=False, clone=not self.reuse_output_buffers)
        return outs


class DRFSISRNet(DRFNet):
    """DRFSISRNet - drop-in for the reference's src/model/nets/drf_sisr_net.py (class at :8, forward at :38-49): the same
    three blocks (identical `state_dict` keys and construction order) iterated `num_steps` times on ONE image,
    `forward(input) -> list of num_steps tensors (N, 1, s*h, s*w)`.  On this path it is the DRFNet engine fed with the
    image repeated `num_steps` times."""

    def __init__(self, in_channels, out_channels, num_steps, num_features, num_groups, upscale_factor):
        super().__init__(in_channels, out_channels, num_features, num_groups, upscale_factor)
        if num_steps < 1:
            raise ValueError(f'The number of the iterations should be positive. Got {num_steps}.')
        self.num_steps = num_steps

    def forward(self, input):
        return super().forward([input] * self.num_steps)
