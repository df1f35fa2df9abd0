"""ORACLE (test infrastructure only - never imported by the product path).

Plain torch.nn.functional fp32 CPU restatement of the reference's EDSRNet
(/root/reference/src/model/nets/edsr_net.py) and of the SISR trainer's loss
(/root/reference/src/runner/trainers/acdc_sisr_trainer.py:27-37: `loss_fn(output, target)` per configured loss).
Pinned against the unmodified reference run in the build container: oracle/make_golden_edsr.py ->
tests/golden/edsr_*.npz (checked by tests/test_edsr.py::test_oracle_matches_reference_golden).
"""
import math

import torch
import torch.nn.functional as F


def up_factors(upscale_factor):
    """_UpBlock (edsr_net.py:60-71): log2(s) x (conv F->4F, PixelShuffle 2) for powers of two; one (conv F->9F,
    PixelShuffle 3) for s = 3."""
    if (math.log(upscale_factor, 2) % 1) == 0:
        return [2] * int(math.log(upscale_factor, 2))
    if upscale_factor == 3:
        return [3]
    raise NotImplementedError


def edsr_forward(sd, x, num_resblocks, upscale_factor, res_scale=0.1):
    """EDSRNet.forward (edsr_net.py:35-39) from a state dict with the reference's keys."""
    head = F.conv2d(x, sd['head.0.weight'], sd['head.0.bias'], padding=1)                         # :29
    y = head
    for i in range(num_resblocks):                                                                # _ResBlock :42-58
        t = F.relu(F.conv2d(y, sd[f'body.{i}.body.conv1.weight'], sd[f'body.{i}.body.conv1.bias'], padding=1))
        res = F.conv2d(t, sd[f'body.{i}.body.conv2.weight'], sd[f'body.{i}.body.conv2.bias'], padding=1) * res_scale
        y = res + y
    y = F.conv2d(y, sd['body.conv.weight'], sd['body.conv.bias'], padding=1) + head               # :31, :37
    for k, r in enumerate(up_factors(upscale_factor)):                                            # _UpBlock
        y = F.pixel_shuffle(F.conv2d(y, sd[f'tail.0.conv{k + 1}.weight'], sd[f'tail.0.conv{k + 1}.bias'], padding=1), r)
    return F.conv2d(y, sd['tail.conv.weight'], sd['tail.conv.bias'], padding=1)                   # :33


def edsr_loss_and_grads(sd, x, target, num_resblocks, upscale_factor, res_scale=0.1):
    """L1 loss (configs/train/edsr_net/exp1_x4.yaml:44-46) and its parameter gradients."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    out = edsr_forward(params, x, num_resblocks, upscale_factor, res_scale)
    loss = F.l1_loss(out, target)
    loss.backward()
    return out.detach(), loss.detach(), {k: p.grad for k, p in params.items()}
