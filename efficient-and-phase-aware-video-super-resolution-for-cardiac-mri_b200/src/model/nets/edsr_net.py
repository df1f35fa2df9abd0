"""EDSRNet - drop-in for the reference's src/model/nets/edsr_net.py (class at :8, forward at :35).

Same constructor, same `forward(input)` contract ((N, 1, h, w) -> (N, 1, s*h, s*w)), same `state_dict` keys / shapes
(`head.0`, `body.{i}.body.conv{1,2}`, `body.conv`, `tail.0.conv{k}`, `tail.conv`) and the same construction order, so
`torch.manual_seed(s)` + construction yields the reference's weights and reference checkpoints load with strict=True.
The sub-modules only hold the fp32 master parameters; the arithmetic runs on the RefineNet conv core (tcgen05
implicit-GEMM kernels behind include/pvsr.h) through pvsr.edsr_engine.EDSREngine.  No CPU / PyTorch fallback.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from .base_net import BaseNet


def _conv(c_in, c_out):
    return nn.Conv2d(c_in, c_out, kernel_size=3, padding=1)


class _ParamResBlock(nn.Module):
    """Parameter holder of _ResBlock (edsr_net.py:42-58): `body.conv1`, `body.conv2` (ReLU between them)."""

    def __init__(self, num_features):
        super().__init__()
        self.body = nn.Sequential(OrderedDict([('conv1', _conv(num_features, num_features)), ('relu1', nn.ReLU()),
                                               ('conv2', _conv(num_features, num_features))]))


class EDSRNet(BaseNet):
    """
    Args:
        in_channels (int): The input channels (1 on this path).
        out_channels (int): The output channels (1 on this path).
        num_resblocks (int): The number of the resblocks.
        num_features (int): The number of the internal feature maps (a multiple of 64, at most 256).
        upscale_factor (int): The upscale factor (2, 3, 4 or 8).
        res_scale (float): The residual scaling factor of the resblocks. Default: `0.1`.
    """

    def __init__(self, in_channels, out_channels, num_resblocks, num_features, upscale_factor, res_scale=0.1):
        super().__init__()
        from pvsr.edsr_engine import up_factors
        factors = up_factors(upscale_factor)          # NotImplementedError for unsupported factors, as the reference
        if in_channels != 1 or out_channels != 1:
            raise ValueError('The B200 path implements the single-channel cine-MRI configuration '
                             f'(in_channels = out_channels = 1). Got {in_channels}, {out_channels}.')
        if num_features % 64 != 0 or not 64 <= num_features <= 256:
            raise ValueError(f'The B200 path needs num_features in (64, 128, 192, 256). Got {num_features}.')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_resblocks, self.num_features = num_resblocks, num_features
        self.upscale_factor, self.res_scale = upscale_factor, res_scale

        self.head = nn.Sequential(_conv(in_channels, num_features))
        self.body = nn.Sequential(*[_ParamResBlock(num_features) for _ in range(num_resblocks)])
        self.body.add_module('conv', _conv(num_features, num_features))
        up = nn.Sequential()
        for k, r in enumerate(factors):
            up.add_module(f'conv{k + 1}', _conv(num_features, r * r * num_features))
            up.add_module(f'deconv{k + 1}', nn.PixelShuffle(r))
        self.tail = nn.Sequential(up)
        self.tail.add_module('conv', _conv(num_features, out_channels))
        self.reuse_output_buffers = False
        self._engine = None

    @property
    def engine(self):
        if self._engine is None:
            from pvsr.edsr_engine import EDSREngine
            self._engine = EDSREngine(self)
        return self._engine

    def forward(self, input):
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            from pvsr.edsr_engine import edsr_train_forward
            return edsr_train_forward(self, input)
        out, _ = self.engine.forward(input, train=False, clone=not self.reuse_output_buffers)
        return out
