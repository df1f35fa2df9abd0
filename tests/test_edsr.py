"""EDSRNet on the RefineNet conv core: the oracle against the reference's golden vectors (CPU), the
module contract (CPU), and forward / gradient parity of the CUDA path against the oracle and the golden vectors (GPU).

Tolerances: the reference itself, run under torch.autocast(bfloat16) against its own fp32 run (32 blocks x 256
features, probed in the build container), drifts by max-abs 4.0e-3 / rel-L2 7.1e-3 on the output and by rel-L2 up to
9.1e-2 / cosine 0.9958 on parameter gradients.  Gates: output max-abs <= 2e-2 and rel-L2 <= 1.5e-2, loss rel <= 2e-3,
gradients rel-L2 <= 0.12 and cosine >= 0.99 per tensor.
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN

CASES = sorted(os.path.basename(p)[len("edsr_"):-4] for p in glob.glob(os.path.join(GOLDEN, "edsr_*.npz")))


def _load(name):
    z = np.load(os.path.join(GOLDEN, f"edsr_{name}.npz"), allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def _net(kwargs):
    from src.model.nets import EDSRNet
    torch.manual_seed(0)
    return EDSRNet(**kwargs)


def test_golden_cases_exist():
    assert len(CASES) >= 5


@pytest.mark.parametrize("name", CASES)
def test_module_reproduces_reference_parameters(name):
    """Same seed + construction -> the reference's weights (state_dict keys, shapes and checksums)."""
    z, meta = _load(name)
    net = _net(meta["kwargs"])
    sd = net.state_dict()
    assert list(sd.keys()) == list(meta["params"].keys())
    for k, (shape, s, sa) in meta["params"].items():
        assert list(sd[k].shape) == shape
        assert float(sd[k].double().sum()) == pytest.approx(s, rel=1e-9, abs=1e-9)
        assert float(sd[k].double().abs().sum()) == pytest.approx(sa, rel=1e-9)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    from oracle import edsr_oracle as O
    z, meta = _load(name)
    kw = meta["kwargs"]
    sd = _net(kw).state_dict()
    x, target = torch.from_numpy(z["input"]), torch.from_numpy(z["target"])
    out, loss, grads = O.edsr_loss_and_grads(sd, x, target, kw["num_resblocks"], kw["upscale_factor"], kw["res_scale"])
    assert (out - torch.from_numpy(z["output"])).abs().max().item() <= 2e-6
    assert abs(loss.item() - float(z["loss"])) <= 1e-6
    for k, (norm, total) in meta["grads"].items():
        assert float(grads[k].double().norm()) == pytest.approx(norm, rel=1e-4)
        want = torch.from_numpy(z["grad::" + k])
        got = grads[k] if grads[k].dim() == 1 else grads[k].reshape(-1)[::meta["stride"]]
        assert (got - want).abs().max().item() <= 1e-6 + 1e-4 * want.abs().max().item()


def test_constructor_contract():
    from src.model.nets import EDSRNet
    with pytest.raises(NotImplementedError):
        EDSRNet(1, 1, 2, 64, 5)
    with pytest.raises(ValueError):
        EDSRNet(3, 3, 2, 64, 4)
    with pytest.raises(ValueError):
        EDSRNet(1, 1, 2, 96, 4)
    net = EDSRNet(1, 1, 32, 256, 4, res_scale=0.1)
    assert sum(p.numel() for p in net.parameters()) == 43080705      # the reference's exp1_x4 model
    from pvsr.lib import PvsrError
    with pytest.raises(PvsrError):
        net.eval()(torch.zeros(1, 1, 8, 8))                          # no CPU fallback


# ------------------------------------------------------------------------------------------------ GPU
def _check_out(out, want):
    d = (out - want)
    assert d.abs().max().item() <= 2e-2, d.abs().max().item()
    assert (d.norm() / want.norm()).item() <= 1.5e-2, (d.norm() / want.norm()).item()


def _check_grads(net, z, meta, grads=None):
    worst_rel, worst_cos = 0.0, 1.0
    for k, p in net.named_parameters():
        g = (p.grad if grads is None else grads[k]).detach().float().cpu()
        norm, _ = meta["grads"][k]
        want = torch.from_numpy(z["grad::" + k])
        got = g if g.dim() == 1 else g.reshape(-1)[::meta["stride"]]
        assert float(g.double().norm()) == pytest.approx(norm, rel=0.12), k
        if want.numel() >= 8 and want.norm() > 0:
            rel = ((got - want).norm() / want.norm()).item()
            cos = (got.double() @ want.double() / (got.double().norm() * want.double().norm())).item()
            worst_rel, worst_cos = max(worst_rel, rel), min(worst_cos, cos)
            assert rel <= 0.12 and cos >= 0.99, (k, rel, cos)
    return worst_rel, worst_cos


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("graph", [False, True])
def test_forward_matches_golden(name, graph, pvsr_lib):
    z, meta = _load(name)
    net = _net(meta["kwargs"]).cuda().eval()
    net.engine.use_graph = graph
    x = torch.from_numpy(z["input"]).cuda()
    want = torch.from_numpy(z["output"])
    with torch.no_grad():
        for _ in range(3 if graph else 1):       # eager, eager + capture, replay
            out = net(x)
    assert out.shape == want.shape and out.dtype == torch.float32
    _check_out(out.cpu(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_autograd_path_matches_golden(name, pvsr_lib):
    """The reference trainer's own sequence: loss_fn(net(x), target).backward() through the autograd bridge."""
    z, meta = _load(name)
    net = _net(meta["kwargs"]).cuda().train()
    x, target = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    for _ in range(3):                            # third pass replays the captured graphs
        net.zero_grad()
        out = net(x)
        loss = torch.nn.L1Loss()(out, target)
        loss.backward()
    _check_out(out.detach().cpu(), torch.from_numpy(z["output"]))
    assert loss.item() == pytest.approx(float(z["loss"]), rel=2e-3)
    _check_grads(net, z, meta)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["x4_f256_r3", "x2_f128_r2"])
def test_fused_step_matches_golden_and_adam(name, pvsr_lib):
    """Fused path: forward + L1 + backward without autograd into the flat gradient buffer, then FusedAdam."""
    from pvsr.optim import FusedAdam
    z, meta = _load(name)
    net = _net(meta["kwargs"]).cuda().train()
    opt = FusedAdam.for_net(net, lr=1e-4)
    ref = _net(meta["kwargs"]).cuda()
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-4)
    x, target = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    loss, out = net.engine.loss_and_grads(x, target)
    assert loss.item() == pytest.approx(float(z["loss"]), rel=2e-3)
    _check_out(out.detach().cpu(), torch.from_numpy(z["output"]))
    _check_grads(net, z, meta)
    for p, q in zip(net.parameters(), ref.parameters()):
        q.grad = p.grad.detach().clone()
    opt.step()
    ref_opt.step()
    for (k, p), q in zip(net.named_parameters(), ref.parameters()):
        assert (p - q).abs().max().item() <= 2e-6, k
    # two more steps: replayed graphs, re-packed weights; the loss must go down on the same batch
    for _ in range(3):
        loss2, _ = net.engine.loss_and_grads(x, target)
        opt.step()
    assert loss2.item() < loss.item()


@pytest.mark.gpu
def test_full_size_edsr_against_oracle(pvsr_lib):
    """The reference's exp1_x4 model (32 blocks x 256 features, 43 M parameters) on one 54x63 ACDC-shaped frame."""
    from oracle import edsr_oracle as O
    kw = dict(in_channels=1, out_channels=1, num_resblocks=32, num_features=256, upscale_factor=4, res_scale=0.1)
    net = _net(kw)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 1, 54, 63, generator=g)
    with torch.no_grad():
        want = O.edsr_forward(sd, x, 32, 4, 0.1)
        out = net.cuda().eval()(x.cuda())
    assert out.shape == (1, 1, 216, 252)
    _check_out(out.cpu(), want)
    from src.model.metrics import PSNR
    from src.utils import denormalize
    a, b = denormalize(out.cpu(), 'acdc'), denormalize(want, 'acdc')
    assert PSNR()(a, b).item() > 45.0          # SR frames agree to within a grey level almost everywhere


# ------------------------------------------------------------------------------------------------ Bicubic baseline
def _bicubic_ref(x, s):
    """The reference's Bicubic.forward (bicubic.py:15-19), executed by torch on the CPU."""
    return torch.nn.Upsample(scale_factor=s, mode='bicubic', align_corners=True)(x)


def test_bicubic_contract():
    from src.model.nets import Bicubic
    from pvsr.lib import PvsrError
    net = Bicubic(upscale_factor=4)
    assert len(net.state_dict()) == 0
    with pytest.raises(PvsrError):
        net(torch.zeros(1, 1, 4, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,s", [((2, 1, 54, 63), 4), ((1, 1, 72, 84), 3), ((3, 1, 108, 126), 2), ((1, 2, 5, 1), 4),
                                     ((1, 1, 1, 7), 2)])
def test_bicubic_matches_torch(shape, s, pvsr_lib):
    from src.model.nets import Bicubic
    g = torch.Generator().manual_seed(11)
    x = torch.randn(*shape, generator=g)
    want = _bicubic_ref(x, s)
    got = Bicubic(s).cuda()(x.cuda()).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("kw,shape", [
    (dict(num_resblocks=1, num_features=192, upscale_factor=2), (2, 1, 10, 14)),     # BN = 64 x 3 column tiles
    (dict(num_resblocks=2, num_features=64, upscale_factor=8), (1, 1, 6, 5)),        # three conv + PixelShuffle(2) stages
    (dict(num_resblocks=0, num_features=128, upscale_factor=3), (3, 1, 7, 9)),       # no residual block at all
])
def test_other_widths_and_factors_against_oracle(kw, shape, pvsr_lib):
    """Configurations without a golden file: the CUDA path against the (golden-pinned) oracle, forward and gradients."""
    from oracle import edsr_oracle as O
    full = dict(in_channels=1, out_channels=1, res_scale=0.1, **kw)
    net = _net(full)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    s = kw["upscale_factor"]
    g = torch.Generator().manual_seed(17)
    x = torch.randn(*shape, generator=g)
    target = torch.randn(shape[0], 1, shape[2] * s, shape[3] * s, generator=g)
    want, loss, grads = O.edsr_loss_and_grads(sd, x, target, kw["num_resblocks"], s, 0.1)
    net = net.cuda().train()
    out = net(x.cuda())
    got_loss = torch.nn.L1Loss()(out, target.cuda())
    got_loss.backward()
    _check_out(out.detach().cpu(), want)
    assert got_loss.item() == pytest.approx(loss.item(), rel=2e-3)
    for k, p in net.named_parameters():
        a, b = p.grad.cpu(), grads[k]
        rel = ((a - b).norm() / b.norm()).item()
        assert rel <= 0.12, (k, rel)
