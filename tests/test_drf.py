"""DRFNet on the RefineNet conv core (reference src/model/nets/drf_net.py): the oracle against the
reference's golden vectors (CPU), the module contract (CPU), the operand tables of the projection units against
torch's ConvTranspose2d / strided Conv2d (CPU), and forward / gradient parity of the CUDA path against the golden
vectors written by the UNMODIFIED reference (GPU).

Tolerances as tests/test_edsr.py (bf16 operands, fp32 accumulation): output max-abs <= 2e-2 and rel-L2 <= 1.5e-2,
loss rel <= 2e-3, gradients rel-L2 <= 0.12 and cosine >= 0.99 per tensor (the reference under bf16 autocast drifts
by up to 9.1e-2 / 0.9958 from its own fp32 run).
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN

CASES = sorted(os.path.basename(p)[len("drfnet_"):-4] for p in glob.glob(os.path.join(GOLDEN, "drfnet_*.npz")))


def _load(name):
    z = np.load(os.path.join(GOLDEN, f"drfnet_{name}.npz"), allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def _net(kwargs):
    from src.model.nets import DRFNet
    torch.manual_seed(0)
    return DRFNet(**kwargs)


def test_golden_cases_exist():
    assert len(CASES) >= 5


@pytest.mark.parametrize("name", CASES)
def test_module_reproduces_reference_parameters(name):
    """Same seed + construction -> the reference's weights (state_dict keys, shapes and checksums)."""
    z, meta = _load(name)
    sd = _net(meta["kwargs"]).state_dict()
    assert list(sd.keys()) == list(meta["params"].keys())
    for k, (shape, s, sa) in meta["params"].items():
        assert list(sd[k].shape) == shape
        assert float(sd[k].double().sum()) == pytest.approx(s, rel=1e-9, abs=1e-9)
        assert float(sd[k].double().abs().sum()) == pytest.approx(sa, rel=1e-9)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    from oracle import drf_oracle as O
    z, meta = _load(name)
    kw = meta["kwargs"]
    sd = _net(kw).state_dict()
    inputs = [torch.from_numpy(x) for x in z["inputs"]]
    targets = [torch.from_numpy(x) for x in z["targets"]]
    outs, loss, grads = O.drf_loss_and_grads(sd, inputs, targets, kw["num_groups"], kw["upscale_factor"])
    for o, w in zip(outs, z["outputs"]):
        assert (o - torch.from_numpy(w)).abs().max().item() <= 2e-6
    assert abs(loss.item() - float(z["loss"])) <= 1e-6
    for k, (norm, total) in meta["grads"].items():
        assert float(grads[k].double().norm()) == pytest.approx(norm, rel=1e-4)
        want = torch.from_numpy(z["grad::" + k])
        got = grads[k] if grads[k].dim() == 1 else grads[k].reshape(-1)[::meta["stride"]]
        assert (got - want).abs().max().item() <= 1e-6 + 1e-4 * want.abs().max().item()


@pytest.mark.parametrize("name", CASES[:2])
def test_slope_probe_reproduces_reference_gradients(name):
    z, meta = _load(name)
    for k, (ref, scale) in _slope_terms(z, meta).items():
        assert ref == pytest.approx(meta["grads"][k][1], rel=1e-3, abs=1e-9), k
        assert scale >= abs(ref)


def _apply_table(idx, w):
    """Packed operand [K blocks, n_total, 64] (fp64) from a gather table and a parameter tensor."""
    flat = np.concatenate([w.reshape(-1), [0.0]])
    return flat[idx]                                        # idx == -1 picks the appended zero


def _conv3x3_by_table(x, op, kb):
    """x [C_in = 64 kb, h, w] -> [n_total, h, w]: the implicit GEMM the kernel runs (taps raster order, zero padding)."""
    c, h, w = x.shape
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1)))
    out = np.zeros((op.shape[1], h, w))
    for tap in range(9):
        dy, dx = tap // 3, tap % 3
        for cb in range(kb):
            blk = op[tap * kb + cb]                         # [n_total, 64]
            out += np.einsum('nc,chw->nhw', blk, xp[cb * 64:(cb + 1) * 64, dy:dy + h, dx:dx + w])
    return out


@pytest.mark.parametrize("s", [2, 3, 4, 8])
def test_projection_tables_equal_torch_convs(s):
    """The phase-stacked 3x3 forms of ConvTranspose2d(k, s, p) and Conv2d(k, s, p) (drf_net.py:69-88), table by table."""
    from pvsr.drf_engine import PROJECTION, table_expand, table_reduce
    k, _, p = PROJECTION[s]
    F, h, w, P = 64, 4, 5, s * s
    g = torch.Generator().manual_seed(s)
    x = torch.randn(1, F, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(F, F, k, k, generator=g, dtype=torch.float64)
    # deconv: F -> P*F at the LR grid, column q*F + c = HR pixel (s*y + q // s, s*x + q % s)
    want = torch.nn.functional.conv_transpose2d(x, wt, stride=s, padding=p)[0].numpy()            # [F, s*h, s*w]
    got = _conv3x3_by_table(x[0].numpy(), _apply_table(table_expand(F, k, s, p), wt.numpy()), 1)
    got = got.reshape(s, s, F, h, w).transpose(2, 3, 0, 4, 1).reshape(F, s * h, s * w)
    assert np.abs(got - want).max() <= 1e-10
    # strided conv on the phase-stacked HR map
    hr = torch.randn(1, F, s * h, s * w, generator=g, dtype=torch.float64)
    want = torch.nn.functional.conv2d(hr, wt, stride=s, padding=p)[0].numpy()                     # [F, h, w]
    stacked = hr[0].numpy().reshape(F, h, s, w, s).transpose(2, 4, 0, 1, 3).reshape(P * F, h, w)
    got = _conv3x3_by_table(stacked, _apply_table(table_reduce(F, k, s, p), wt.numpy()), P)
    assert np.abs(got - want).max() <= 1e-10


@pytest.mark.parametrize("s", [2, 3, 4, 8])
def test_projection_pairs_cover_every_kernel_tap_once(s):
    """The sparse weight-gradient form: the non-zero (tap, phase) blocks of the phase-stacked convs are exactly the
    k*k taps of the k x k parameter, each once, and they are exactly the non-zero blocks of the dense tables."""
    from pvsr.drf_engine import PROJECTION, projection_pairs, table_expand, table_reduce
    k, _, p = PROJECTION[s]
    F, P = 64, s * s
    for kind, table in (("expand", table_expand(F, k, s, p)), ("reduce", table_reduce(F, k, s, p))):
        pairs = projection_pairs(kind, k, s, p)
        assert len(pairs) == k * k and sorted((ky, kx) for _, _, ky, kx in pairs) == [(a, b) for a in range(k) for b in range(k)]
        nz = set()
        if kind == "expand":            # table [tap][q * F + ch][c]
            t = table.reshape(9, P, F, 64)
            for tap in range(9):
                for q in range(P):
                    if (t[tap, q] >= 0).any():
                        assert (t[tap, q] >= 0).all()
                        nz.add((tap, q))
        else:                           # table [tap * P + q][col][c]
            t = table.reshape(9, P, F, 64)
            for tap in range(9):
                for q in range(P):
                    if (t[tap, q] >= 0).any():
                        assert (t[tap, q] >= 0).all()
                        nz.add((tap, q))
        assert nz == {(tap, q) for tap, q, _, _ in pairs}


def test_pointwise_tables():
    from pvsr.drf_engine import table_pointwise, table_pointwise_T
    F, n_src = 64, 3
    w = np.random.RandomState(0).randn(F, n_src * F)
    op = _apply_table(table_pointwise(F, n_src, F), w)                     # [n_src, F, 64]
    for j in range(n_src):
        assert np.array_equal(op[j], w[:, j * F:(j + 1) * F])
        assert np.array_equal(_apply_table(table_pointwise_T(F, n_src, F, j), w)[0], w[:, j * F:(j + 1) * F].T)


def test_constructor_contract():
    from src.model.nets import DRFNet
    with pytest.raises(ValueError):
        DRFNet(1, 1, 64, 3, 5)
    with pytest.raises(ValueError):
        DRFNet(3, 3, 64, 3, 4)
    with pytest.raises(ValueError):
        DRFNet(1, 1, 96, 3, 4)
    net = DRFNet(1, 1, 64, 6, 4)
    assert sum(p.numel() for p in net.parameters()) == 3658907
    from pvsr.lib import PvsrError
    with pytest.raises(PvsrError):
        net.eval()([torch.zeros(1, 1, 8, 8)])                        # no CPU fallback


# ------------------------------------------------------------------------------------------------ DRFSISRNet
SISR_CASES = sorted(os.path.basename(p)[len("drfsisr_"):-4] for p in glob.glob(os.path.join(GOLDEN, "drfsisr_*.npz")))


def _load_sisr(name):
    z = np.load(os.path.join(GOLDEN, f"drfsisr_{name}.npz"), allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def _sisr_net(kwargs):
    from src.model.nets import DRFSISRNet
    torch.manual_seed(0)
    return DRFSISRNet(**kwargs)


@pytest.mark.parametrize("name", SISR_CASES)
def test_drfsisr_module_and_oracle_match_reference_golden(name):
    """DRFSISRNet (drf_sisr_net.py): same seed -> the reference's weights; the oracle (the DRFNet restatement fed with the
    image repeated num_steps times) reproduces the reference's outputs, SRFB-trainer loss and gradients."""
    from oracle import drf_oracle as O
    z, meta = _load_sisr(name)
    kw = meta["kwargs"]
    sd = _sisr_net(kw).state_dict()
    assert list(sd.keys()) == list(meta["params"].keys())
    for k, (shape, s, sa) in meta["params"].items():
        assert list(sd[k].shape) == shape and float(sd[k].double().sum()) == pytest.approx(s, rel=1e-9, abs=1e-9)
    x, target = torch.from_numpy(z["input"]), torch.from_numpy(z["target"])
    S = kw["num_steps"]
    outs, loss, grads = O.drf_loss_and_grads(sd, [x] * S, [target] * S, kw["num_groups"], kw["upscale_factor"])
    for o, w in zip(outs, z["outputs"]):
        assert (o - torch.from_numpy(w)).abs().max().item() <= 2e-6
    assert abs(loss.item() - float(z["loss"])) <= 1e-6
    for k, (norm, total) in meta["grads"].items():
        assert float(grads[k].double().norm()) == pytest.approx(norm, rel=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SISR_CASES)
def test_drfsisr_matches_golden(name, pvsr_lib):
    """Inference and the SRFB trainer's sequence (acdc_sisr_srfb_trainer.py:22-26) through the autograd bridge."""
    z, meta = _load_sisr(name)
    kw = meta["kwargs"]
    x, target = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    net = _sisr_net(kw).cuda().eval()
    with torch.no_grad():
        outs = net(x)
    assert len(outs) == kw["num_steps"]
    for o, w in zip(outs, z["outputs"]):
        _check_out(o.cpu(), torch.from_numpy(w))
    net.train()
    for _ in range(3):
        net.zero_grad()
        outs = net(x)
        loss = torch.stack([torch.nn.L1Loss()(o, target) for o in outs]).mean()
        loss.backward()
    assert loss.item() == pytest.approx(float(z["loss"]), rel=2e-3)
    for k, p in net.named_parameters():
        if p.numel() > 1:
            norm, _ = meta["grads"][k]
            want = torch.from_numpy(z["grad::" + k])
            g = p.grad.detach().float().cpu()
            got = g if g.dim() == 1 else g.reshape(-1)[::meta["stride"]]
            assert float(g.double().norm()) == pytest.approx(norm, rel=0.12), k
            if want.numel() >= 8:
                assert ((got - want).norm() / want.norm()).item() <= 0.12, k


# ------------------------------------------------------------------------------------------------ GPU
def _check_out(out, want):
    d = (out - want)
    assert d.abs().max().item() <= 2e-2, d.abs().max().item()
    assert (d.norm() / want.norm()).item() <= 1.5e-2, (d.norm() / want.norm()).item()


def _slope_terms(z, meta):
    """Cancellation scale of every PReLU slope gradient (a scalar = the sum of T*n*h*w*C terms g * min(z, 0) of both
    signs): the CPU oracle's (sum, absolute sum) of the terms.  On the fixtures |sum| / abs-sum is 1e-4 .. 1e-2, so the
    bf16 path is gated at 0.12 |ref| + 2.5e-3 abs-sum (bf16 eps = 3.9e-3; measured <= 1.4e-3 abs-sum)."""
    from oracle import drf_oracle as O
    kw = meta["kwargs"]
    sd = _net(kw).state_dict()
    return O.slope_gradient_terms(sd, [torch.from_numpy(x) for x in z["inputs"]],
                                  [torch.from_numpy(x) for x in z["targets"]], kw["num_groups"], kw["upscale_factor"])


def _check_grads(net, z, meta):
    terms = _slope_terms(z, meta)
    for k, p in net.named_parameters():
        g = p.grad.detach().float().cpu()
        norm, total = meta["grads"][k]
        if k in terms:
            ref, scale = terms[k]
            assert ref == pytest.approx(total, rel=1e-3, abs=1e-9), k          # the probe reproduces the reference's value
            assert abs(float(g) - total) <= 0.12 * abs(total) + 2.5e-3 * scale, (k, float(g), total, scale)
            continue
        want = torch.from_numpy(z["grad::" + k])
        got = g if g.dim() == 1 else g.reshape(-1)[::meta["stride"]]
        assert float(g.double().norm()) == pytest.approx(norm, rel=0.12), k
        if want.numel() >= 8 and want.norm() > 0:
            rel = ((got - want).norm() / want.norm()).item()
            cos = (got.double() @ want.double() / (got.double().norm() * want.double().norm())).item()
            assert rel <= 0.12 and cos >= 0.99, (k, rel, cos)


@pytest.mark.gpu
def test_prelu_stream_kernels(pvsr_lib):
    from pvsr import lib as L
    g = torch.Generator(device="cuda").manual_seed(1)
    z = torch.randn(3, 5, 7, 64, generator=g, device="cuda").to(torch.bfloat16)
    gy = torch.randn(3, 5, 7, 64, generator=g, device="cuda").to(torch.bfloat16)
    a = torch.tensor([0.2], device="cuda")
    y = torch.empty_like(z)
    L.check(pvsr_lib.pvsr_prelu_fwd_bf16(L.ptr(z), L.ptr(a), L.ptr(y), z.numel(), L.current_stream()), "prelu_fwd")
    assert torch.equal(y, torch.nn.functional.prelu(z.float(), a).to(torch.bfloat16))
    zf = z.float().requires_grad_(True)
    af = a.clone().requires_grad_(True)
    (torch.nn.functional.prelu(zf, af) * gy.float()).sum().backward()
    da = torch.zeros(1, device="cuda")
    dz = torch.empty_like(z)
    L.check(pvsr_lib.pvsr_prelu_bwd_bf16(L.ptr(gy), L.ptr(z), L.ptr(a), L.ptr(dz), L.ptr(da), z.numel(), L.current_stream()),
            "prelu_bwd")
    assert torch.equal(dz, zf.grad.to(torch.bfloat16))
    assert da.item() == pytest.approx(af.grad.item(), rel=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("graph", [False, True])
def test_forward_matches_golden(name, graph, pvsr_lib):
    z, meta = _load(name)
    net = _net(meta["kwargs"]).cuda().eval()
    net.engine.use_graph = graph
    inputs = [torch.from_numpy(x).cuda() for x in z["inputs"]]
    with torch.no_grad():
        for _ in range(3 if graph else 1):       # eager, eager + capture, replay
            outs = net(inputs)
    assert len(outs) == len(inputs)
    for o, w in zip(outs, z["outputs"]):
        assert o.shape == w.shape and o.dtype == torch.float32
        _check_out(o.cpu(), torch.from_numpy(w))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_autograd_path_matches_golden(name, pvsr_lib):
    """The reference trainer's own sequence (acdc_vsr_trainer.py:40-47) through the autograd bridge."""
    z, meta = _load(name)
    net = _net(meta["kwargs"]).cuda().train()
    inputs = [torch.from_numpy(x).cuda() for x in z["inputs"]]
    targets = [torch.from_numpy(x).cuda() for x in z["targets"]]
    for _ in range(3):                            # third pass replays the captured graphs
        net.zero_grad()
        outs = net(inputs)
        loss = torch.stack([torch.nn.L1Loss()(o, t) for o, t in zip(outs, targets)]).mean()
        loss.backward()
    for o, w in zip(outs, z["outputs"]):
        _check_out(o.detach().cpu(), torch.from_numpy(w))
    assert loss.item() == pytest.approx(float(z["loss"]), rel=2e-3)
    _check_grads(net, z, meta)


@pytest.mark.gpu
def test_fused_step_matches_golden(pvsr_lib):
    z, meta = _load("x4_g3")
    net = _net(meta["kwargs"]).cuda().train()
    inputs = [torch.from_numpy(x).cuda() for x in z["inputs"]]
    targets = [torch.from_numpy(x).cuda() for x in z["targets"]]
    for _ in range(3):
        loss, outs = net.engine.loss_and_grads(inputs, targets)
    assert loss.item() == pytest.approx(float(z["loss"]), rel=2e-3)
    _check_grads(net, z, meta)


@pytest.mark.gpu
def test_acdc_shaped_sequence_against_oracle(pvsr_lib):
    """x4, 6 groups, 5 frames of one 54x63 ACDC-shaped sequence against the CPU oracle."""
    from oracle import drf_oracle as O
    kw = dict(in_channels=1, out_channels=1, num_features=64, num_groups=6, upscale_factor=4)
    net = _net(kw)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(9)
    inputs = [torch.randn(1, 1, 54, 63, generator=g) for _ in range(5)]
    with torch.no_grad():
        want = O.drf_forward(sd, inputs, 6, 4)
        outs = net.cuda().eval()([x.cuda() for x in inputs])
    for o, w in zip(outs, want):
        assert o.shape == (1, 1, 216, 252)
        _check_out(o.cpu(), w)
