"""CPU tests (no GPU): the oracle restatement and the drop-in module against fixtures produced by the
UNMODIFIED reference (oracle/make_golden.py).  These pin the oracle before any CUDA result is trusted."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, build_net, golden_cases, load_golden, oracle_kwargs
from oracle import refinenet_oracle as O

REF_KEYS_X4 = 26


@pytest.mark.parametrize("name", golden_cases())
def test_dropin_init_matches_reference(name):
    """Same seed -> same parameters as the reference module (keys, shapes, checksums)."""
    z, meta = load_golden(name)
    net = build_net(meta["kwargs"])
    sd = dict(net.named_parameters())
    assert list(sd.keys()) == list(meta["params"].keys())
    for k, (shape, s, a) in meta["params"].items():
        assert list(sd[k].shape) == shape, k
        assert abs(float(sd[k].detach().double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), k
        assert abs(float(sd[k].detach().double().abs().sum()) - a) <= 1e-9 * max(1.0, a), k
    if meta["kwargs"]["upscale_factor"] == 4 and meta["kwargs"].get("positional_encoding") and len(meta["kwargs"]["num_features"]) == 3:
        assert len(net.state_dict()) == REF_KEYS_X4


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_forward_matches_reference(name):
    z, meta = load_golden(name)
    net = build_net(meta["kwargs"])
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    inputs = [torch.from_numpy(x) for x in z["inputs"]]
    pos = torch.from_numpy(z["pos"])
    with torch.no_grad():
        out = O.refinenet_forward(sd, inputs, pos, **oracle_kwargs(meta["kwargs"]))
    got = torch.stack([torch.stack(o) for o in out]).numpy()
    ref = z["outputs_eval"]
    assert got.shape == ref.shape
    # fp32 CPU, same operators in a different association order: 2e-6 absolute on values of O(0.1)
    assert np.abs(got - ref).max() <= 2e-6


@pytest.mark.parametrize("name", ["x4_pos", "x3_pos", "x4_nopos", "x4_nomem"])
def test_oracle_loss_and_grads_match_reference(name):
    """Training-mode gradient rule (detach at non-grad frames == the reference's no_grad blocks)."""
    z, meta = load_golden(name)
    net = build_net(meta["kwargs"])
    sd = dict(net.named_parameters())
    for k in list(sd):
        sd[k].grad = None
    inputs = [torch.from_numpy(x) for x in z["inputs"]]
    pos = torch.from_numpy(z["pos"])
    targets = [torch.from_numpy(t) for t in z["targets"]]
    out = O.refinenet_forward(sd, inputs, pos, train=True, **oracle_kwargs(meta["kwargs"]))
    loss = O.trainer_loss(out, targets, training=True)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) <= 1e-5
    for k, g in meta["grads"].items():
        if g is None:
            assert sd[k].grad is None or float(sd[k].grad.abs().sum()) == 0.0, k   # dead PReLU of the refine block
            continue
        norm, _ = g
        got = float(sd[k].grad.double().norm())
        assert abs(got - norm) <= 1e-4 * max(norm, 1e-6) + 1e-7, (k, got, norm)
        if "grad::" + k in z.files:
            assert np.allclose(sd[k].grad.numpy(), z["grad::" + k], rtol=1e-3, atol=1e-6), k


def test_known_answers():
    """The KAT values recorded in SURVEY.md section 8c (N=2, L=19, 8x8) reproduced by the oracle."""
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        kat = json.load(f)
    assert abs(kat["x4"]["out_sum"] - (-51.776825)) < 1e-4 and abs(kat["x4"]["loss"] - 4.227020) < 1e-5
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=3, update_memory=True,
              num_updated_frames=6, refine_window_size=5, upscale_factor=4, positional_encoding=True)
    net = build_net(kw)
    assert sum(p.numel() for p in net.parameters()) == kat["x4"]["n_params"] == 2890993
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    inputs = [torch.randn(2, 1, 8, 8, generator=g) for _ in range(19)]
    pos = torch.randn(2, 19, 1, generator=g)
    targets = [torch.randn(2, 1, 32, 32, generator=g) for _ in range(7)]
    with torch.no_grad():
        out = O.refinenet_forward(sd, inputs, pos, train=True, **oracle_kwargs(kw))
        loss = O.trainer_loss(out, targets)
    assert abs(float(out[-1][0].sum()) - kat["x4"]["out_sum"]) < 1e-4
    assert np.allclose([float(v) for v in out[-1][0][0, 0, 0, :4]], kat["x4"]["out_first4"], atol=1e-6)
    assert abs(float(loss) - kat["x4"]["loss"]) < 1e-5


def test_metrics_restatement():
    """PSNR/SSIM/denormalize follow src/model/metrics.py and src/utils.py (spot values computed by hand)."""
    g = torch.Generator().manual_seed(0)
    a = torch.rand(2, 1, 32, 32, generator=g)
    d = O.denormalize((a * 255 - 54.089) / 48.084)
    assert torch.all(d == d.round()) and d.min() >= 0 and d.max() <= 255
    assert torch.allclose(d, (a * 255).round(), atol=1.0)
    b = (d + 3).clamp(0, 255)
    mse = ((d - b) ** 2).mean(dim=(1, 2, 3))
    assert abs(float(O.psnr(d, b)) - float((10 * torch.log10(255 ** 2 / (mse + 1e-10))).mean())) < 1e-5
    assert abs(float(O.ssim(d, d)) - 1.0) < 1e-5
    k = O.ssim_kernel()
    assert abs(float(k.sum()) - 1.0) < 1e-6 and k.shape == (1, 1, 11, 11)
    # the reference's non-standard Gaussian: exp(-((x-mu)/(2 sigma))^2), i.e. variance 2*sigma^2
    assert abs(float(k[0, 0, 5, 6] / k[0, 0, 5, 5]) - float(np.exp(-(1 / 3.0) ** 2))) < 1e-6


def test_positional_code_and_window():
    p = O.positional_code(30, 11)
    assert p.shape == (30,) and abs(p[0] - 1.0) < 1e-7 and abs(p[11] + 1.0) < 1e-6
    frames = list(range(30))
    win = O.circular_window(frames, 30, 6)
    assert len(win) == 42 and win[:6] == [24, 25, 26, 27, 28, 29] and win[6] == 0 and win[-6:] == [0, 1, 2, 3, 4, 5]
