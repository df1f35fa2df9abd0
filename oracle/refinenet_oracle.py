"""ORACLE - TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU fp32 restatement (plain torch.nn.functional) of the reference hot path
`/root/reference/src/model/nets/refine_net.py` (RefineNet.forward + its blocks), of the trainer's
multi-stage loss, and of the metrics the reference evaluates with.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this file.

Parity pin: PINNED.  The restatement is checked against outputs of the *unmodified* reference model
run in the build container (oracle/make_golden.py imports /root/reference and writes
tests/golden/*.npz; tests/test_oracle.py compares).  The reference ships no tests/golden vectors of its
own, and its arithmetic lives in torch (conda pytorch=1.3.0 in env.yml:168; torch 2.11
here, same operator semantics), so the live reference is the pin.

The functions take a `state_dict` with the reference's 26 keys (refine_net.py:18-59).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ blocks
def in_block(sd, x):
    """_InBlock: conv3x3 + PReLU (refine_net.py:188-192)."""
    y = F.conv2d(x, sd["in_block.conv.weight"], sd["in_block.conv.bias"], padding=1)
    return F.prelu(y, sd["in_block.prelu.weight"])


def lstm_cell(sd, prefix, x, h, c, memory=True):
    """ConvLSTMCell.forward (refine_net.py:247-267); gate order i, f, o, g (:258)."""
    combined = torch.cat([x, h], dim=1) if memory else torch.cat([x, x], dim=1)  # :253 / :255
    cc = F.conv2d(combined, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], padding=1)
    hd = cc.shape[1] // 4
    cc_i, cc_f, cc_o, cc_g = torch.split(cc, hd, dim=1)
    i, f, o, g = torch.sigmoid(cc_i), torch.sigmoid(cc_f), torch.sigmoid(cc_o), torch.tanh(cc_g)
    c_next = f * c + i * g
    h_next = o * torch.tanh(c_next)
    return h_next, c_next


def out_block(sd, x, scale):
    """_OutBlock (refine_net.py:194-205)."""
    if scale in (2, 4, 8):
        n = int(math.log2(scale))
        for i in range(n):
            x = F.conv2d(x, sd[f"out_block.conv{i + 1}.weight"], sd[f"out_block.conv{i + 1}.bias"], padding=1)
            x = F.pixel_shuffle(x, 2)
        return F.conv2d(x, sd[f"out_block.conv{n + 1}.weight"], sd[f"out_block.conv{n + 1}.bias"], padding=1)
    x = F.conv2d(x, sd["out_block.conv1.weight"], sd["out_block.conv1.bias"], padding=1)
    x = F.pixel_shuffle(x, 3)
    return F.conv2d(x, sd["out_block.conv2.weight"], sd["out_block.conv2.bias"], padding=1)


def refine_body(sd, feat, positional_encoding):
    """_RefineBlock.body (refine_net.py:147-155): conv1 -> conv2 with NO activation (the PReLU is registered on
    the block, never applied), or a single 1x1 conv without positional encoding."""
    if positional_encoding:
        y = F.conv2d(feat, sd["refine_block.body.conv1.weight"], sd["refine_block.body.conv1.bias"], padding=1)
        return F.conv2d(y, sd["refine_block.body.conv2.weight"], sd["refine_block.body.conv2.bias"], padding=1)
    return F.conv2d(feat, sd["refine_block.body.conv1.weight"], sd["refine_block.body.conv1.bias"])


# ------------------------------------------------------------------------------------------------ forward
def refinenet_forward(sd, inputs, pos_codes, *, num_stages=3, num_updated_frames=6, refine_window_size=5,
                      upscale_factor=4, positional_encoding=True, memory=True, num_layers=3, train=False):
    """RefineNet.forward (refine_net.py:61-135) in the unified per-frame indexing of SURVEY.md Appendix A.

    inputs: list of L tensors (N, 1, h, w); pos_codes: (N, L, 1).  Returns a tuple of 3*num_stages lists of T
    tensors (N, 1, s*h, s*w), order per stage [forward head, backward head, fused head] (:100-113).
    train=True reproduces the reference's no_grad blocks by detaching everything produced at a non-grad frame
    (:74-79, 82-93, 179-183).
    """
    U, Wn, S = num_updated_frames, refine_window_size, num_stages
    L = len(inputs)
    if U <= 0:
        raise IndexError("num_updated_frames must be > 0 (the reference crashes on inputs[0:-0], :66)")
    T = L - 2 * U
    half = Wn // 2
    N, _, h, w = inputs[0].shape
    is_grad = lambda j: U <= j < L - U
    cut = (lambda t, j: t if (is_grad(j) or not train) else t.detach())

    x = [cut(in_block(sd, inputs[j]), j) for j in range(L)]  # :66-67, 74-79
    outputs = []
    for _ in range(S):
        zeros = lambda: torch.zeros(N, 64, h, w, dtype=x[0].dtype)
        hf, hb = [None] * L, [None] * L
        for direction, prefix, store in (("f", "forward_lstm_block", hf), ("b", "backward_lstm_block", hb)):
            state = [(zeros(), zeros()) for _ in range(num_layers)]  # :71-72
            order = range(L) if direction == "f" else range(L - 1, -1, -1)
            for j in order:  # :82-93
                cur = x[j]
                for l in range(num_layers):
                    hh, cc = lstm_cell(sd, f"{prefix}.cell_list.{l}", cur, state[l][0], state[l][1], memory)
                    hh, cc = cut(hh, j), cut(cc, j)
                    state[l] = (hh, cc)
                    cur = hh
                store[j] = cur
        # refine block (:157-185)
        r = [None] * L
        for j in range(half, L - half):
            chans = []
            for d in range(-half, half + 1):
                chans += [hf[j + d], hb[j + d]]
                if positional_encoding:
                    chans.append(pos_codes[:, j + d].view(N, 1, 1, 1).expand(N, 1, h, w))
            r[j] = cut(refine_body(sd, torch.cat(chans, dim=1), positional_encoding), j)
        # heads (:100-113)
        outputs.append([out_block(sd, x[j] + hf[j], upscale_factor) for j in range(U, L - U)])
        outputs.append([out_block(sd, x[j] + hb[j], upscale_factor) for j in range(U, L - U)])
        outputs.append([out_block(sd, x[j] + r[j], upscale_factor) for j in range(U, L - U)])
        # feature updates (:118-133)
        if S > 1:
            x = [x[j] + (hf[j] if j < half else hb[j] if j >= L - half else r[j]) for j in range(L)]
    return tuple(outputs)


# ------------------------------------------------------------------------------------------------ loss / metrics
def trainer_loss(outputs, targets, training=True):
    """AcdcVSRRefineNetTrainer._compute_losses with nn.L1Loss, weight 1
    (src/runner/trainers/acdc_vsr_refinenet_trainer.py:75-101)."""
    if training:
        losses = []
        for i, outs in enumerate(outputs):
            discount = float(np.power(0.5, (len(outputs) // 3 - i // 3 - 1)))
            losses.append(torch.stack([F.l1_loss(o, t) * discount for o, t in zip(outs, targets)]).mean())
        return torch.stack(losses).sum()
    return torch.stack([F.l1_loss(o, t) for o, t in zip(outputs[-1], targets)]).mean()


def denormalize(imgs, dataset="acdc"):
    """src/utils.py:1-20."""
    mean, std = {"acdc": (54.089, 48.084), "dsb15": (51.193, 52.671)}[dataset]
    return (imgs.clone() * std + mean).round().clamp(0, 255)


def psnr(output, target, max_value=255):
    """src/model/metrics.py:20-36 (size_average=True)."""
    dims = list(range(1, output.dim()))
    mse = F.mse_loss(output, target, reduction="none").mean(dims)
    return (10 * torch.log10(max_value ** 2 / (mse + 1e-10))).mean()


def ssim_kernel():
    """src/model/metrics.py:66-82: 11x11 'Gaussian' exp(-((x-mu)/(2 sigma))^2), sigma 1.5, normalised."""
    g = torch.arange(11, dtype=torch.float32)
    k1 = 1 / (1.5 * math.sqrt(2 * math.pi)) * torch.exp(-((g - 5) / (2 * 1.5)) ** 2)
    k = k1[:, None] * k1[None, :]
    return (k / k.sum()).view(1, 1, 11, 11)


def ssim(output, target, value_range=255):
    """src/model/metrics.py:86-113 (dim=2, channels=1, size_average=True; valid convolution)."""
    wgt = ssim_kernel().to(output)
    c1, c2 = (0.01 * value_range) ** 2, (0.03 * value_range) ** 2
    mu1, mu2 = F.conv2d(output, wgt), F.conv2d(target, wgt)
    s1 = F.conv2d(output * output, wgt) - mu1.pow(2)
    s2 = F.conv2d(target * target, wgt) - mu2.pow(2)
    s12 = F.conv2d(output * target, wgt) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2.0 * s12 + c2)) / ((mu1.pow(2) + mu2.pow(2) + c1) * (s1 + s2 + c2))
    return m.mean()


# ------------------------------------------------------------------------------------------------ synthetic data
def positional_code(T, end_systole):
    """src/gen_positional_encoding.py:35-38: two half cosines split at the end-systole frame."""
    y1 = np.cos(np.linspace(0, np.pi, end_systole, endpoint=False))
    y2 = np.cos(np.linspace(np.pi, np.pi * 2, T - end_systole, endpoint=False))
    return np.concatenate((y1, y2)).astype(np.float32)


def circular_window(frames, T, U):
    """Test-mode slicing of AcdcVSRRefineNetDataset.__getitem__ (acdc_vsr_refinenet_dataset.py:74-87):
    the cycle is tiled three times and frames [T-U, 2T+U) are taken."""
    tiled = list(frames) * 3
    return tiled[T - U:2 * T + U]


def init_state_dict(upscale_factor=4, positional_encoding=True, window=5, num_layers=3, seed=0):
    """Random-init parameters with the reference's 26 keys/shapes and torch's default Conv2d init
    (kaiming_uniform(a=sqrt(5)) + uniform bias), PReLU 0.2 (refine_net.py:18-59,192).  Values are NOT identical
    to constructing the reference module under the same seed; use tests/golden for that."""
    g = torch.Generator().manual_seed(seed)

    def conv(co, ci, k):
        bound = 1.0 / math.sqrt(ci * k * k)
        wgt = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound
        b = (torch.rand(co, generator=g) * 2 - 1) * bound
        return wgt, b

    sd = {}
    sd["in_block.conv.weight"], sd["in_block.conv.bias"] = conv(64, 1, 3)
    sd["in_block.prelu.weight"] = torch.full((1,), 0.2)
    for d in ("forward", "backward"):
        for l in range(num_layers):
            p = f"{d}_lstm_block.cell_list.{l}.conv"
            sd[p + ".weight"], sd[p + ".bias"] = conv(256, 128, 3)
    if positional_encoding:
        cin = window * 129
        sd["refine_block.body.conv1.weight"], sd["refine_block.body.conv1.bias"] = conv(129, cin, 3)
        sd["refine_block.body.conv2.weight"], sd["refine_block.body.conv2.bias"] = conv(64, 129, 3)
    else:
        sd["refine_block.body.conv1.weight"], sd["refine_block.body.conv1.bias"] = conv(64, window * 128, 1)
    sd["refine_block.prelu.weight"] = torch.full((1,), 0.2)
    if upscale_factor in (2, 4, 8):
        n = int(math.log2(upscale_factor))
        for i in range(n):
            sd[f"out_block.conv{i + 1}.weight"], sd[f"out_block.conv{i + 1}.bias"] = conv(256, 64, 3)
        sd[f"out_block.conv{n + 1}.weight"], sd[f"out_block.conv{n + 1}.bias"] = conv(1, 64, 3)
    else:
        sd["out_block.conv1.weight"], sd["out_block.conv1.bias"] = conv(576, 64, 3)
        sd["out_block.conv2.weight"], sd["out_block.conv2.bias"] = conv(1, 64, 3)
    return sd
