"""RefineNet - drop-in for the reference's src/model/nets/refine_net.py (class at :10, forward at :61).

Same constructor, same `forward(inputs, pos_codes)` contract, same 26 `state_dict` keys/shapes and the same
default initialisation order (so `torch.manual_seed(s)` + construction yields the reference's weights, and
reference checkpoints load with strict=True).  The sub-modules below only *hold* the fp32 master parameters;
the arithmetic of forward runs in hand-written sm_100a kernels behind the C ABI (include/pvsr.h) through
pvsr.engine.RefineNetEngine.  There is no PyTorch / CPU fallback: without a CUDA device and libpvsr.so the
forward raises.
"""
import math

import torch
import torch.nn as nn

from .base_net import BaseNet


class _ParamConvLSTMCell(nn.Module):
    """Parameter holder with the reference layout of ConvLSTMCell (refine_net.py:208-245)."""

    def __init__(self, input_dim, hidden_dim, memory):
        super().__init__()
        self.input_dim, self.hidden_dim, self.memory = input_dim, hidden_dim, memory
        in_ch = input_dim + hidden_dim if memory else input_dim * 2
        self.conv = nn.Conv2d(in_ch, 4 * hidden_dim, kernel_size=(3, 3), padding=(1, 1), bias=True)


class _ParamConvLSTM(nn.Module):
    """Parameter holder of _ConvLSTM (refine_net.py:274-302): `cell_list.{i}.conv.{weight,bias}`."""

    def __init__(self, input_dim, hidden_dim, num_layers, memory):
        super().__init__()
        cells = []
        for i in range(num_layers):
            cur = input_dim if i == 0 else hidden_dim[i - 1]
            cells.append(_ParamConvLSTMCell(cur, hidden_dim[i], memory))
        self.cell_list = nn.ModuleList(cells)


class _ParamRefineBlock(nn.Module):
    """Parameter holder of _RefineBlock (refine_net.py:138-155), including the never-applied PReLU that the
    reference registers on the block (its weight is part of state_dict and never receives a gradient)."""

    def __init__(self, in_channels, num_features, num_frames, positional_encoding):
        super().__init__()
        self.body = nn.Sequential()
        if positional_encoding:
            self.body.add_module('conv1', nn.Conv2d(in_channels, in_channels // num_frames, kernel_size=3, padding=1))
            self.add_module('prelu', nn.PReLU(num_parameters=1, init=0.2))
            self.body.add_module('conv2', nn.Conv2d(in_channels // num_frames, num_features, kernel_size=3, padding=1))
            self.add_module('prelu', nn.PReLU(num_parameters=1, init=0.2))
        else:
            self.body.add_module('conv1', nn.Conv2d(in_channels, num_features, kernel_size=1))
            self.add_module('prelu', nn.PReLU(num_parameters=1, init=0.2))


class _ParamInBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
        self.prelu = nn.PReLU(num_parameters=1, init=0.2)


class _ParamOutBlock(nn.Module):
    """Parameter holder of _OutBlock (refine_net.py:194-205): conv{1..} with PixelShuffle between them."""

    def __init__(self, in_channels, out_channels, upscale_factor):
        super().__init__()
        if (math.log(upscale_factor, 2) % 1) == 0:
            n = int(math.log(upscale_factor, 2))
            for i in range(n):
                self.add_module(f'conv{i + 1}', nn.Conv2d(in_channels, 4 * in_channels, kernel_size=3, padding=1))
            self.add_module(f'conv{n + 1}', nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1))
            self.num_ps_convs = n
        elif upscale_factor == 3:
            self.add_module('conv1', nn.Conv2d(in_channels, 9 * in_channels, kernel_size=3, padding=1))
            self.add_module('conv2', nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1))
            self.num_ps_convs = 1


class RefineNet(BaseNet):
    """
    Args:
        in_channels (int): The input channels (1 on this path).
        out_channels (int): The output channels (1 on this path).
        num_features (list of int): The number of the internal feature maps (64 each).
        upscale_factor (int): The upscale factor (2, 3, 4 or 8).
    Extra attributes (not in the reference):
        only_last_head (bool): in eval/no_grad mode return just the last output list (what the predictor
            consumes through `[-1]`) and skip the other 3*num_stages-1 heads.  Default False = reference behaviour.
        reuse_output_buffers (bool): return views of the engine's output buffer instead of fresh tensors.
    """

    def __init__(self, in_channels, out_channels, num_features, num_stages=1, refine_window_size=5, upscale_factor=4,
                 update_memory=False, num_updated_frames=0, memory=True, positional_encoding=False):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_features = num_features
        self.num_stages = num_stages
        self.refine_window_size = refine_window_size
        self.upscale_factor = upscale_factor
        self.update_memory = update_memory
        self.num_updated_frames = num_updated_frames
        self.memory = memory
        self.positional_encoding = positional_encoding

        if upscale_factor not in [2, 3, 4, 8]:
            raise ValueError(f'The upscale factor should be 2, 3, 4 or 8. Got {upscale_factor}.')
        if update_memory == False and num_updated_frames != 0:
            raise ValueError('The \"update_memory\" is not activated!')
        if in_channels != 1 or out_channels != 1:
            raise ValueError('The B200 path implements the single-channel cine-MRI configuration '
                             f'(in_channels = out_channels = 1). Got {in_channels}, {out_channels}.')
        if any(f != 64 for f in num_features) or not 1 <= len(num_features) <= 3:
            raise ValueError(f'The B200 path implements 1-3 ConvLSTM layers of 64 features. Got {num_features}.')

        num_feature = num_features[0]
        self.in_block = _ParamInBlock(in_channels, num_feature)
        self.forward_lstm_block = _ParamConvLSTM(num_feature, num_features, len(num_features), memory)
        self.backward_lstm_block = _ParamConvLSTM(num_feature, num_features, len(num_features), memory)
        if positional_encoding:
            refine_in_features = refine_window_size * (num_features[-1] * 2 + 1)
        else:
            refine_in_features = refine_window_size * (num_features[-1] * 2)
        self.refine_block = _ParamRefineBlock(refine_in_features, num_features[-1], refine_window_size,
                                              positional_encoding)
        self.out_block = _ParamOutBlock(num_feature, out_channels, upscale_factor)
        self.num_head_convs = self.out_block.num_ps_convs + 1

        self.only_last_head = False
        self.reuse_output_buffers = False
        self._engine = None

    @property
    def engine(self):
        if self._engine is None:
            from pvsr.engine import RefineNetEngine
            self._engine = RefineNetEngine(self)
        return self._engine

    def forward(self, inputs, pos_codes):
        """inputs: list of L = T + 2U tensors (N, 1, h, w); pos_codes: (N, L, 1).
        Returns a tuple of 3*num_stages lists (forward head, backward head, fused head per stage) of T tensors
        (N, 1, s*h, s*w) - reference refine_net.py:100-113,135."""
        if self.num_updated_frames <= 0:
            # reference: inputs[0:-0] is empty -> IndexError at refine_net.py:71
            raise IndexError('list index out of range (num_updated_frames must be > 0)')
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            from pvsr.autograd import refinenet_train_forward
            return refinenet_train_forward(self, inputs, pos_codes)
        all_heads = not self.only_last_head
        return self.engine.forward(inputs, pos_codes, all_heads=all_heads, clone=not self.reuse_output_buffers)
