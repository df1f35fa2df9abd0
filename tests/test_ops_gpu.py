"""Kernel-level parity (GPU): every sm_100a kernel against a plain PyTorch fp32 reference of the same op.

Inputs and weights are made exactly bf16-representable, so the only differences are the fp32 accumulation
order and the final bf16 rounding of the stored activations.  Tolerances are written at each comparison.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def bf16r(t):
    return t.to(torch.bfloat16).float()


def nhwc(x):  # [n, C, H, W] fp32 -> [n, H, W, C] bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x):  # [n, H, W, C] -> [n, C, H, W] fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _rand(*shape, gen, scale=1.0):
    return bf16r(torch.randn(*shape, generator=gen, device="cuda") * scale)


@pytest.mark.parametrize("n,H,W", [(3, 10, 13), (2, 54, 63), (1, 63, 48), (2, 32, 32), (1, 5, 130)])
def test_conv_store_256(pvsr_lib, n, H, W):
    from pvsr import ops, lib as L
    g = torch.Generator(device="cuda").manual_seed(1)
    x = _rand(n, 64, H, W, gen=g)
    w = _rand(256, 64, 3, 3, gen=g, scale=0.05)
    b = torch.randn(256, generator=g, device="cuda")
    spec = ops._spec(256, 64, 3, 1, [0], 64, 1, 9, 256)
    wp = ops.pack_weight(w, spec)
    out = torch.zeros(n, H, W, 256, dtype=torch.float32, device="cuda")
    ops.conv3x3(nhwc(x), [0], n, wp, 256, bias=ops.pack_bias(b, spec), out_f32=out)
    torch.cuda.synchronize()
    ref = F.conv2d(x, w, b, padding=1)
    assert torch.allclose(nchw(out), ref, atol=2e-4, rtol=1e-4), (nchw(out) - ref).abs().max()


def test_conv_store_64_residual_144in(pvsr_lib):
    """refine conv2 shape: 129 (stored as 144) -> 64 channels, + bias + bf16 residual."""
    from pvsr import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    n, H, W = 4, 54, 63
    x = _rand(n, 129, H, W, gen=g)
    xs = torch.zeros(n, 144, H, W, device="cuda")
    xs[:, :129] = x
    w = _rand(64, 129, 3, 3, gen=g, scale=0.03)
    b = torch.randn(64, generator=g, device="cuda")
    r = _rand(n, 64, H, W, gen=g)
    spec = ops.spec_refine_conv2()
    wp = ops.pack_weight(w, spec)
    out = torch.zeros(n, H, W, 64, dtype=torch.bfloat16, device="cuda")
    ops.conv3x3(nhwc(xs), [0], n, wp, 64, kb_per_src=3, k16_last=1, bias=ops.pack_bias(b, spec), out_bf16=out,
                res=nhwc(r))
    torch.cuda.synchronize()
    ref = F.conv2d(x, w, b, padding=1) + r
    # bf16 output rounding: 2^-9 relative
    assert torch.allclose(nchw(out), ref, atol=1e-2, rtol=8e-3), (nchw(out) - ref).abs().max()
    assert rel_l2(nchw(out), ref) < 3e-3


@pytest.mark.parametrize("B,H,W", [(2, 54, 63), (3, 9, 17)])
def test_refine_conv1_posterm(pvsr_lib, B, H, W):
    """645 -> 129 conv over a 5-frame window of (fwd h | bwd h | pos plane), refine_net.py:166-180."""
    from pvsr import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    Lf, win = 9, 5
    nf = Lf - win + 1
    hf = _rand(Lf * B, 64, H, W, gen=g)
    hb = _rand(Lf * B, 64, H, W, gen=g)
    pos = torch.randn(B, Lf, generator=g, device="cuda")
    w1 = _rand(129, 645, 3, 3, gen=g, scale=0.02)
    b1 = torch.randn(129, generator=g, device="cuda")
    spec = ops.spec_refine_conv1()
    wp = ops.pack_weight(w1, spec)
    table = ops.refine_posterm(w1, b1, pos, nf)
    act = torch.cat([nhwc(hf), nhwc(hb)], dim=0)  # images: [hf frames | hb frames]
    src = []
    for j in range(win):
        src += [j * B, Lf * B + j * B]
    out = torch.zeros(nf * B, H, W, 144, dtype=torch.bfloat16, device="cuda")
    ops.conv3x3(act, src, nf * B, wp, 144, posterm=table, out_bf16=out, out_ch=144, n_store=144)
    torch.cuda.synchronize()
    # reference: build the 645-channel window input exactly like the reference does
    hf5 = hf.view(Lf, B, 64, H, W)
    hb5 = hb.view(Lf, B, 64, H, W)
    refs = []
    for i in range(nf):
        chans = []
        for j in range(win):
            p = pos[:, i + j].view(B, 1, 1, 1).expand(B, 1, H, W)
            chans += [hf5[i + j], hb5[i + j], p]
        refs.append(F.conv2d(torch.cat(chans, dim=1), w1, b1, padding=1))
    ref = torch.stack(refs).view(nf * B, 129, H, W)
    got = nchw(out)
    assert torch.allclose(got[:, :129], ref, atol=2e-2, rtol=8e-3), (got[:, :129] - ref).abs().max()
    assert rel_l2(got[:, :129], ref) < 3e-3
    assert got[:, 129:].abs().max().item() == 0.0


def test_conv1x1_window(pvsr_lib):
    """_RefineBlock without positional encoding: 1x1 conv 640 -> 64 (refine_net.py:154)."""
    from pvsr import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    B, H, W, Lf, win = 2, 20, 23, 7, 5
    nf = Lf - win + 1
    hf = _rand(Lf * B, 64, H, W, gen=g)
    hb = _rand(Lf * B, 64, H, W, gen=g)
    w = _rand(64, 640, 1, 1, gen=g, scale=0.05)
    b = torch.randn(64, generator=g, device="cuda")
    spec = ops.spec_refine_conv1x1()
    wp = ops.pack_weight(w, spec)
    act = torch.cat([nhwc(hf), nhwc(hb)], dim=0)
    src = []
    for j in range(win):
        src += [j * B, Lf * B + j * B]
    out = torch.zeros(nf * B, H, W, 64, dtype=torch.float32, device="cuda")
    ops.conv3x3(act, src, nf * B, wp, 64, taps=1, bias=ops.pack_bias(b, spec), out_f32=out)
    torch.cuda.synchronize()
    hf5, hb5 = hf.view(Lf, B, 64, H, W), hb.view(Lf, B, 64, H, W)
    refs = []
    for i in range(nf):
        chans = []
        for j in range(win):
            chans += [hf5[i + j], hb5[i + j]]
        refs.append(F.conv2d(torch.cat(chans, dim=1), w, b))
    ref = torch.stack(refs).view(nf * B, 64, H, W)
    assert torch.allclose(nchw(out), ref, atol=3e-4, rtol=1e-4), (nchw(out) - ref).abs().max()


@pytest.mark.parametrize("r,H,W", [(2, 54, 63), (3, 24, 28), (2, 7, 9)])
def test_head_conv_pixel_shuffle(pvsr_lib, r, H, W):
    from pvsr import ops, lib as L
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 3
    x = _rand(n, 64, H, W, gen=g)
    w = _rand(64 * r * r, 64, 3, 3, gen=g, scale=0.05)
    b = torch.randn(64 * r * r, generator=g, device="cuda")
    spec = ops.spec_head_ps(r)
    wp = ops.pack_weight(w, spec)
    out = torch.zeros(n, H * r, W * r, 64, dtype=torch.bfloat16, device="cuda")
    bn, nt = (256, 1) if r == 2 else (192, 3)
    ops.conv3x3(nhwc(x), [0], n, wp, bn, epi=L.EPI_PS, bias=ops.pack_bias(b, spec), n_tiles_n=nt, out_bf16=out,
                out_ch=64, ps_r=r)
    torch.cuda.synchronize()
    ref = F.pixel_shuffle(F.conv2d(x, w, b, padding=1), r)
    assert torch.allclose(nchw(out), ref, atol=1e-2, rtol=8e-3), (nchw(out) - ref).abs().max()
    assert rel_l2(nchw(out), ref) < 3e-3


@pytest.mark.parametrize("n,H,W", [(2, 54, 63), (5, 12, 20)])
def test_convlstm_cell(pvsr_lib, n, H, W):
    """Two consecutive ConvLSTMCell steps (refine_net.py:247-267), first with h = c = 0."""
    from pvsr import ops, lib as L
    g = torch.Generator(device="cuda").manual_seed(6)
    x0 = _rand(n, 64, H, W, gen=g)
    x1 = _rand(n, 64, H, W, gen=g)
    w = _rand(256, 128, 3, 3, gen=g, scale=0.04)
    b = torch.randn(256, generator=g, device="cuda") * 0.5
    spec = ops.spec_lstm()
    wp = ops.pack_weight(w, spec)
    bias = ops.pack_bias(b, spec)

    def ref_cell(x, h, c):
        cc = F.conv2d(torch.cat([x, h], dim=1), w, b, padding=1)
        i, f, o, gg = torch.split(cc, 64, dim=1)
        c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        return torch.sigmoid(o) * torch.tanh(c2), c2

    z = torch.zeros(n, 64, H, W, device="cuda")
    h1r, c1r = ref_cell(x0, z, z)
    h1r_b = bf16r(h1r)
    h2r, c2r = ref_cell(x1, h1r_b, c1r)

    # images: [x0 | x1 | h1]
    act = torch.zeros(3 * n, H, W, 64, dtype=torch.bfloat16, device="cuda")
    act[:n] = nhwc(x0)
    act[n:2 * n] = nhwc(x1)
    cst = ops.lstm_state(n, H, W, "cuda")
    gates = torch.zeros(cst.numel() * 4, dtype=torch.bfloat16, device="cuda")
    # step 0: only the x source (h = 0), c_in = None
    ops.conv3x3(act, [0], n, wp, 256, epi=L.EPI_LSTM, bias=bias, c_in=None, c_out=cst, h_out=act[2 * n:],
                gates_out=gates)
    torch.cuda.synchronize()
    h1 = nchw(act[2 * n:])
    c1 = ops.lstm_state_to_nchw(cst, n, H, W)
    assert torch.allclose(c1, c1r, atol=2e-4, rtol=1e-4), (c1 - c1r).abs().max()
    assert torch.allclose(h1, h1r, atol=5e-3, rtol=8e-3), (h1 - h1r).abs().max()
    # step 1: sources x1 and h1 (as produced by the kernel), state updated in place
    ops.conv3x3(act, [n, 2 * n], n, wp, 256, epi=L.EPI_LSTM, bias=bias, c_in=cst, c_out=cst,
                h_out=act[:n])
    torch.cuda.synchronize()
    h2r, c2r = ref_cell(x1, h1, c1)
    c2 = ops.lstm_state_to_nchw(cst, n, H, W)
    h2 = nchw(act[:n])
    assert torch.allclose(c2, c2r, atol=3e-4, rtol=1e-4), (c2 - c2r).abs().max()
    assert torch.allclose(h2, h2r, atol=5e-3, rtol=8e-3), (h2 - h2r).abs().max()


@pytest.mark.parametrize("n,H,W", [(5, 54, 63), (2, 32, 32), (3, 7, 1), (1, 1, 9), (2, 5, 130)])
def test_in_conv_prelu(pvsr_lib, n, H, W):
    """Run-based stencil (csrc/stencil.cuh): full / partial / single-pixel runs of 8, single rows and columns."""
    from pvsr import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(n, H, W, generator=g, device="cuda")
    w = torch.randn(64, 1, 3, 3, generator=g, device="cuda") * 0.3
    b = torch.randn(64, generator=g, device="cuda") * 0.1
    a = torch.tensor([0.2], device="cuda")
    out = ops.in_conv_prelu(x, w, b, a)
    torch.cuda.synchronize()
    ref = F.prelu(F.conv2d(x.unsqueeze(1), w, b, padding=1), a)
    assert torch.allclose(nchw(out), ref, atol=1e-2, rtol=8e-3), (nchw(out) - ref).abs().max()
    assert rel_l2(nchw(out), ref) < 3e-3


@pytest.mark.parametrize("H,W", [(216, 252), (128, 128), (24, 28), (9, 33), (1, 1), (15, 31)])
def test_head_conv_last_and_l1(pvsr_lib, H, W):
    """64 -> 1 conv at HR.  Like every conv of the path it consumes bf16 operands (activations AND weights) and
    accumulates in fp32: exact against a reference built from the bf16-rounded weights, 2^-9-level against fp32."""
    from pvsr import ops
    g = torch.Generator(device="cuda").manual_seed(8)
    n = 3
    x = _rand(n, 64, H, W, gen=g)
    w = torch.randn(1, 64, 3, 3, generator=g, device="cuda") * 0.05
    b = torch.randn(1, generator=g, device="cuda")
    tgt = torch.randn(n, H, W, generator=g, device="cuda")
    part = torch.zeros(n, device="cuda")
    out = ops.head_conv_last(nhwc(x), w, b, target=tgt, l1_partial=part)
    torch.cuda.synchronize()
    ref = F.conv2d(x, bf16r(w), b, padding=1)[:, 0]
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-4), (out - ref).abs().max()
    ref32 = F.conv2d(x, w, b, padding=1)[:, 0]
    assert rel_l2(out, ref32) < 3e-3
    l1 = (ref - tgt).abs().sum(dim=(1, 2))
    assert torch.allclose(part, l1, rtol=1e-4)


def test_add_bf16(pvsr_lib):
    from pvsr import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    a = torch.randn(3, 54, 63, 64, generator=g, device="cuda").to(torch.bfloat16)
    b = torch.randn(3, 54, 63, 64, generator=g, device="cuda").to(torch.bfloat16)
    out = ops.add_bf16(a, b)
    torch.cuda.synchronize()
    assert torch.equal(out, (a.float() + b.float()).to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------ data gradients
def _autograd_dx(fn, x, gy):
    x = x.clone().requires_grad_(True)
    y = fn(x)
    (gx,) = torch.autograd.grad(y, x, gy)
    return gx


def test_dgrad_plain_conv_accumulates(pvsr_lib):
    """dX of a 64->256 conv = conv of dY with the transposed, spatially flipped weights; EPI_GRAD does +=."""
    from pvsr import ops, lib as L
    g = torch.Generator(device="cuda").manual_seed(11)
    n, H, W = 3, 20, 27
    w = _rand(256, 64, 3, 3, gen=g, scale=0.05)
    gy = _rand(n, 256, H, W, gen=g)
    x = torch.randn(n, 64, H, W, generator=g, device="cuda")
    ref = _autograd_dx(lambda t: F.conv2d(t, w, None, padding=1), x, gy)
    spec = ops._spec(256, 64, 3, 1, [0], 256, 4, 9, 64, transpose_flip=1)
    wp = ops.pack_weight(w, spec)
    acc0 = torch.randn(n, H, W, 64, generator=g, device="cuda")
    acc = acc0.clone()
    ops.conv3x3(nhwc(gy), [0], n, wp, 64, epi=L.EPI_GRAD, kb_per_src=4, grad0=acc, n_store=64)
    torch.cuda.synchronize()
    got = (acc - acc0).permute(0, 3, 1, 2)
    assert torch.allclose(got, ref, atol=3e-3, rtol=1e-3), (got - ref).abs().max()


def test_dgrad_lstm_split(pvsr_lib):
    """ConvLSTM gate conv (128 -> 256): d[x | h] in one launch, routed to two gradient tensors."""
    from pvsr import ops, lib as L
    g = torch.Generator(device="cuda").manual_seed(12)
    n, H, W = 2, 32, 32
    w = _rand(256, 128, 3, 3, gen=g, scale=0.04)
    gy = _rand(n, 256, H, W, gen=g)
    xin = torch.randn(n, 128, H, W, generator=g, device="cuda")
    ref = _autograd_dx(lambda t: F.conv2d(t, w, None, padding=1), xin, gy)
    spec = ops._spec(256, 128, 3, 1, [0], 256, 4, 9, 128, transpose_flip=1)
    wp = ops.pack_weight(w, spec)
    d0 = torch.zeros(n, H, W, 64, device="cuda")
    d1 = torch.zeros(n, H, W, 64, device="cuda")
    ops.conv3x3(nhwc(gy), [0], n, wp, 128, epi=L.EPI_GRAD, kb_per_src=4, grad0=d0, grad1=d1, grad_split=1,
                n_store=128)
    torch.cuda.synchronize()
    got = torch.cat([d0, d1], dim=3).permute(0, 3, 1, 2)
    assert torch.allclose(got, ref, atol=3e-3, rtol=1e-3), (got - ref).abs().max()
    # dual (non-split) routing: the same 64 columns added to two tensors
    spec64 = ops._spec(256, 64, 3, 1, [0], 256, 4, 9, 64, transpose_flip=1)
    w64 = _rand(256, 64, 3, 3, gen=g, scale=0.04)
    a, b = torch.zeros(n, H, W, 64, device="cuda"), torch.ones(n, H, W, 64, device="cuda")
    ops.conv3x3(nhwc(gy), [0], n, ops.pack_weight(w64, spec64), 64, epi=L.EPI_GRAD, kb_per_src=4, grad0=a, grad1=b,
                n_store=64)
    torch.cuda.synchronize()
    assert torch.equal(a + 1.0, b)


@pytest.mark.parametrize("r,H,W", [(2, 27, 31), (3, 12, 14), (2, 64, 126)])
def test_dgrad_pixel_shuffle_conv(pvsr_lib, r, H, W):
    """dX of PixelShuffle(conv(x)): the pixel-unshuffle is done by TMA element strides on the HR gradient."""
    from pvsr import ops, lib as L
    g = torch.Generator(device="cuda").manual_seed(13)
    n = 2
    w = _rand(64 * r * r, 64, 3, 3, gen=g, scale=0.05)
    ghr = _rand(n, 64, H * r, W * r, gen=g)
    x = torch.randn(n, 64, H, W, generator=g, device="cuda")
    ref = _autograd_dx(lambda t: F.pixel_shuffle(F.conv2d(t, w, None, padding=1), r), x, ghr)
    spec = ops._spec(64 * r * r, 64, 3, r * r, [0] * (r * r), 64, 1, 9, 64, transpose_flip=1, k_ps_r=r)
    wp = ops.pack_weight(w, spec)
    out = torch.zeros(n, H, W, 64, dtype=torch.bfloat16, device="cuda")
    srcs = [(0, 0, 0, q % r, q // r) for q in range(r * r)]
    ops.conv3x3([(nhwc(ghr), r)], srcs, n, wp, 64, out_bf16=out, out_hw=(H, W))
    torch.cuda.synchronize()
    got = nchw(out)
    assert rel_l2(got, ref) < 3e-3, rel_l2(got, ref)
    assert torch.allclose(got, ref, atol=3e-2, rtol=1e-2), (got - ref).abs().max()


# ------------------------------------------------------------------------------------------------ weight gradients
def _autograd_dw(fn, w, b, gy):
    w = w.clone().requires_grad_(True)
    b = b.clone().requires_grad_(True)
    y = fn(w, b)
    gw, gb = torch.autograd.grad(y, (w, b), gy)
    return gw, gb


def _check_wgrad(got_w, got_b, ref_w, ref_b):
    # bf16 operands (exact products), fp32 accumulation in a different order
    assert rel_l2(got_w, ref_w) < 2e-4, rel_l2(got_w, ref_w)
    assert rel_l2(got_b, ref_b) < 2e-4, rel_l2(got_b, ref_b)


@pytest.mark.parametrize("n,H,W,splits", [(3, 20, 27, 0), (2, 54, 63, 5), (1, 9, 130, 1)])
def test_wgrad_lstm_shape(pvsr_lib, n, H, W, splits):
    """dW, db of the ConvLSTM gate conv: two 64-channel sources [x | h], 256 output columns."""
    from pvsr import ops
    g = torch.Generator(device="cuda").manual_seed(21)
    x = _rand(n, 64, H, W, gen=g)
    h = _rand(n, 64, H, W, gen=g)
    gy = _rand(n, 256, H, W, gen=g)
    w = torch.randn(256, 128, 3, 3, generator=g, device="cuda") * 0.05
    b = torch.zeros(256, device="cuda")
    ref_w, ref_b = _autograd_dw(lambda ww, bb: F.conv2d(torch.cat([x, h], 1), ww, bb, padding=1), w, b, gy)
    act = torch.cat([nhwc(x), nhwc(h)], dim=0)
    dy = nhwc(gy)
    srcs = [(0, 0, 0, 0, 0), (0, n, 0, 0, 0)]
    dys = [(1, 0, 64 * c, 0, 0) for c in range(4)]
    dwp, dbp = ops.conv3x3_wgrad([(act, 1), (dy, 1)], srcs, dys, n, (H, W), 256, n_splits=splits)
    spec = ops.spec_lstm()
    gw, gb = torch.zeros_like(w), torch.zeros_like(b)
    ops.scatter_add(gw, spec, dwp, gb, dbp)
    torch.cuda.synchronize()
    _check_wgrad(gw, gb, ref_w, ref_b)


def test_wgrad_refine_conv2_shape(pvsr_lib):
    """129 (stored 144) -> 64 conv: three channel blocks per tap, the last one mostly padding."""
    from pvsr import ops
    g = torch.Generator(device="cuda").manual_seed(22)
    n, H, W = 3, 24, 31
    x = _rand(n, 129, H, W, gen=g)
    xs = torch.zeros(n, 144, H, W, device="cuda")
    xs[:, :129] = x
    gy = _rand(n, 64, H, W, gen=g)
    w = torch.randn(64, 129, 3, 3, generator=g, device="cuda") * 0.05
    b = torch.zeros(64, device="cuda")
    ref_w, ref_b = _autograd_dw(lambda ww, bb: F.conv2d(x, ww, bb, padding=1), w, b, gy)
    dwp, dbp = ops.conv3x3_wgrad([(nhwc(xs), 1), (nhwc(gy), 1)], [(0, 0, 0, 0, 0)], [(1, 0, 0, 0, 0)], n, (H, W), 64,
                                 kb_per_src=3)
    gw, gb = torch.zeros_like(w), torch.zeros_like(b)
    ops.scatter_add(gw, ops.spec_refine_conv2(), dwp, gb, dbp)
    torch.cuda.synchronize()
    _check_wgrad(gw, gb, ref_w, ref_b)


def test_wgrad_refine_conv1_shape(pvsr_lib):
    """10 sources (5-frame window of fwd/bwd hidden maps) against a 129 (stored 144) channel gradient."""
    from pvsr import ops
    g = torch.Generator(device="cuda").manual_seed(23)
    B, H, W, Lf, win = 2, 16, 19, 7, 5
    nf = Lf - win + 1
    hf = _rand(Lf * B, 64, H, W, gen=g)
    hb = _rand(Lf * B, 64, H, W, gen=g)
    gy = _rand(nf * B, 129, H, W, gen=g)
    gys = torch.zeros(nf * B, 144, H, W, device="cuda")
    gys[:, :129] = gy
    w = torch.randn(129, 645, 3, 3, generator=g, device="cuda") * 0.02
    b = torch.zeros(129, device="cuda")
    hf5, hb5 = hf.view(Lf, B, 64, H, W), hb.view(Lf, B, 64, H, W)

    def fwd(ww, bb):
        outs = []
        for i in range(nf):
            chans = []
            for j in range(win):
                chans += [hf5[i + j], hb5[i + j], torch.zeros(B, 1, H, W, device="cuda")]
            outs.append(F.conv2d(torch.cat(chans, 1), ww, bb, padding=1))
        return torch.stack(outs).view(nf * B, 129, H, W)

    ref_w, ref_b = _autograd_dw(fwd, w, b, gy)
    act = torch.cat([nhwc(hf), nhwc(hb)], dim=0)
    srcs = []
    for j in range(win):
        srcs += [(0, j * B, 0, 0, 0), (0, Lf * B + j * B, 0, 0, 0)]
    dys = [(1, 0, 64 * c, 0, 0) for c in range(3)]
    dwp, dbp = ops.conv3x3_wgrad([(act, 1), (nhwc(gys), 1)], srcs, dys, nf * B, (H, W), 192)
    spec = ops.spec_refine_conv1()
    spec.n_total = 192
    gw, gb = torch.zeros_like(w), torch.zeros_like(b)
    ops.scatter_add(gw, spec, dwp, gb, dbp)
    torch.cuda.synchronize()
    pos_ch = [129 * j + 128 for j in range(win)]
    keep = [c for c in range(645) if c not in pos_ch]
    _check_wgrad(gw[:, keep], gb, ref_w[:, keep], ref_b)


@pytest.mark.parametrize("r,H,W", [(2, 27, 31), (3, 12, 14)])
def test_wgrad_pixel_shuffle_conv(pvsr_lib, r, H, W):
    from pvsr import ops
    g = torch.Generator(device="cuda").manual_seed(24)
    n = 2
    x = _rand(n, 64, H, W, gen=g)
    ghr = _rand(n, 64, H * r, W * r, gen=g)
    w = torch.randn(64 * r * r, 64, 3, 3, generator=g, device="cuda") * 0.05
    b = torch.zeros(64 * r * r, device="cuda")
    ref_w, ref_b = _autograd_dw(lambda ww, bb: F.pixel_shuffle(F.conv2d(x, ww, bb, padding=1), r), w, b, ghr)
    dys = [(1, 0, 0, q % r, q // r) for q in range(r * r)]
    dwp, dbp = ops.conv3x3_wgrad([(nhwc(x), 1), (nhwc(ghr), r)], [(0, 0, 0, 0, 0)], dys, n, (H, W), 64 * r * r)
    gw, gb = torch.zeros_like(w), torch.zeros_like(b)
    ops.scatter_add(gw, ops.spec_head_ps(r), dwp, gb, dbp)
    torch.cuda.synchronize()
    _check_wgrad(gw, gb, ref_w, ref_b)
