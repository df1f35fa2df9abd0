"""ORACLE (test infrastructure only - never imported by the product path).

Plain torch.nn.functional fp32 CPU restatement of the reference's DRFNet
(/root/reference/src/model/nets/drf_net.py) and of the VSR trainer's loss
(/root/reference/src/runner/trainers/acdc_vsr_trainer.py:40-43,83-94: one `loss_fn(output_t, target_t)` per frame, their
mean per configured loss, weighted sum).
Pinned against the unmodified reference run in the build container: oracle/make_golden_drf.py ->
tests/golden/drfnet_*.npz (checked by tests/test_drf.py::test_oracle_matches_reference_golden).
"""
import math

import torch
import torch.nn.functional as F

# (kernel_size, stride, padding) of the projection units, drf_net.py:69-76
PROJECTION = {2: (6, 2, 2), 3: (7, 3, 2), 4: (8, 4, 2), 8: (12, 8, 2)}


def up_factors(upscale_factor):
    """_OutBlock (drf_net.py:136-147)."""
    if (math.log(upscale_factor, 2) % 1) == 0:
        return [2] * int(math.log(upscale_factor, 2))
    if upscale_factor == 3:
        return [3]
    raise ValueError(f'The upscale factor should be 2, 3, 4 or 8. Got {upscale_factor}.')


# Optional probe (tests only): PROBE = {} makes every PReLU use a per-element copy A of its slope, kept in PROBE[name];
# after backward, A.grad holds the individual terms g * min(z, 0) whose SUM is the slope gradient - their absolute sum is
# the cancellation scale against which a bf16 path's error on that scalar has to be judged.
PROBE = None


def _prelu(sd, name, z):
    a = sd[name + '.weight']
    if PROBE is None:
        return F.prelu(z, a)
    A = a.detach().expand_as(z).clone().requires_grad_(True)
    PROBE.setdefault(name + '.weight', []).append(A)
    return torch.where(z > 0, z, A * z)


def _cp(sd, name, x, act=None, **kw):
    y = F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], **kw)
    return _prelu(sd, act, y) if act else y


def f_block(sd, x, hidden, num_groups, upscale_factor):
    """_FBlock.forward (drf_net.py:118-133)."""
    k, s, p = PROJECTION[upscale_factor]
    P = 'f_block.'
    lr = _cp(sd, P + 'in_block.conv', torch.cat([x, hidden], dim=1), P + 'in_block.prelu')               # :119-120
    lr_list, hr_list = [lr], []
    for i in range(num_groups):                                                                          # :123-129
        u, d = f'{P}up_blocks.{i}.', f'{P}down_blocks.{i}.'
        cat_lr = torch.cat(lr_list, dim=1)
        if i == 0:                                                                                       # :78-88
            hr = _prelu(sd, u + 'prelu',
                        F.conv_transpose2d(cat_lr, sd[u + 'deconv.weight'], sd[u + 'deconv.bias'], stride=s, padding=p))
        else:                                                                                            # :90-95
            m = _cp(sd, u + 'conv1', cat_lr, u + 'prelu1')
            hr = _prelu(sd, u + 'prelu2',
                        F.conv_transpose2d(m, sd[u + 'deconv2.weight'], sd[u + 'deconv2.bias'], stride=s, padding=p))
        hr_list.append(hr)
        cat_hr = torch.cat(hr_list, dim=1)
        if i == 0:
            lr = _cp(sd, d + 'conv', cat_hr, d + 'prelu', stride=s, padding=p)
        else:                                                                                            # :97-102
            m = _cp(sd, d + 'conv1', cat_hr, d + 'prelu1')
            lr = _cp(sd, d + 'conv2', m, d + 'prelu2', stride=s, padding=p)
        lr_list.append(lr)
    return _cp(sd, P + 'out_block.conv', torch.cat(lr_list[1:], dim=1), P + 'out_block.prelu')           # :131-133


def out_block(sd, x, upscale_factor):
    """_OutBlock (drf_net.py:136-147)."""
    fs = up_factors(upscale_factor)
    for i, r in enumerate(fs):
        x = F.pixel_shuffle(_cp(sd, f'out_block.conv{i + 1}', x, padding=1), r)
    return _cp(sd, f'out_block.conv{len(fs) + 1}', x, padding=1)


def drf_forward(sd, inputs, num_groups, upscale_factor):
    """DRFNet.forward (drf_net.py:38-49) from a state dict with the reference's keys."""
    outputs, hidden = [], None
    for i, x in enumerate(inputs):
        a = _cp(sd, 'in_block.conv1', x, 'in_block.prelu1', padding=1)                                   # _InBlock :52-58
        a = _cp(sd, 'in_block.conv2', a, 'in_block.prelu2')
        if i == 0:
            hidden = a                                                                                   # :42-43
        f = f_block(sd, a, hidden, num_groups, upscale_factor)
        hidden = f                                                                                       # :45
        outputs.append(out_block(sd, a + f, upscale_factor))                                             # :46-47
    return outputs


def vsr_loss(outputs, targets):
    """AcdcVSRTrainer with one L1Loss of weight 1 (acdc_vsr_trainer.py:40-43,83-94): mean over frames of L1(out_t, tgt_t)."""
    return torch.stack([F.l1_loss(o, t) for o, t in zip(outputs, targets)]).mean()


def drf_loss_and_grads(sd, inputs, targets, num_groups, upscale_factor):
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    outs = drf_forward(params, inputs, num_groups, upscale_factor)
    loss = vsr_loss(outs, targets)
    loss.backward()
    return [o.detach() for o in outs], loss.detach(), {k: p.grad for k, p in params.items()}


def slope_gradient_terms(sd, inputs, targets, num_groups, upscale_factor):
    """{PReLU slope name: (sum, absolute sum) of the per-element terms g * min(z, 0) of its gradient}."""
    global PROBE
    PROBE = {}
    try:
        outs = drf_forward({k: v.detach() for k, v in sd.items()}, inputs, num_groups, upscale_factor)
        vsr_loss(outs, targets).backward()
        return {k: (sum(float(A.grad.double().sum()) for A in v), sum(float(A.grad.double().abs().sum()) for A in v))
                for k, v in PROBE.items()}
    finally:
        PROBE = None
