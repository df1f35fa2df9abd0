"""Model-level parity (GPU): the drop-in RefineNet (CUDA kernels through the C ABI) against
 (a) fixtures produced by the unmodified reference (tests/golden, small shapes, all 3*S output lists), and
 (b) the pinned CPU oracle at the full ACDCSR shape (LR 54x63, T=30, U=6), incl. PSNR/SSIM parity.

Tolerances (SURVEY.md section 8c error budget for bf16 operands with fp32 accumulation/state):
  outputs max-abs <= 2e-2 and rel-L2 <= 1.5e-2;  PSNR delta <= 0.01 dB, SSIM delta <= 1e-4.
"""
import numpy as np
import pytest
import torch

from helpers import build_net, golden_cases, load_golden, oracle_kwargs

pytestmark = pytest.mark.gpu

MAX_ABS, REL_L2 = 2e-2, 1.5e-2


def _run(net, inputs, pos):
    with torch.no_grad():
        out = net([x.cuda() for x in inputs], pos.cuda())
    torch.cuda.synchronize()
    return out


def _stack(out):
    return torch.stack([torch.stack(list(o)) for o in out]).float().cpu()


@pytest.mark.parametrize("name", golden_cases())
def test_forward_matches_reference_fixture(pvsr_lib, name):
    z, meta = load_golden(name)
    net = build_net(meta["kwargs"]).cuda().eval()
    inputs = [torch.from_numpy(x) for x in z["inputs"]]
    pos = torch.from_numpy(z["pos"])
    out = _run(net, inputs, pos)
    assert isinstance(out, tuple) and len(out) == 3 * meta["kwargs"]["num_stages"]
    assert all(isinstance(o, list) and len(o) == meta["T"] for o in out)
    got, ref = _stack(out), torch.from_numpy(z["outputs_eval"])
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert err <= MAX_ABS and rel <= REL_L2, (name, err, rel)
    # per-list check: every head of every stage, not just the last one
    for l in range(got.shape[0]):
        rl = ((got[l] - ref[l]).norm() / ref[l].norm()).item()
        assert rl <= REL_L2, (name, l, rl)


def test_last_head_fast_path_and_graph_replay(pvsr_lib):
    z, meta = load_golden("x4_pos")
    net = build_net(meta["kwargs"]).cuda().eval()
    inputs = [torch.from_numpy(x) for x in z["inputs"]]
    pos = torch.from_numpy(z["pos"])
    full = _stack(_run(net, inputs, pos))
    net.only_last_head = True
    a = _stack(_run(net, inputs, pos))          # first call: eager + capture
    b = _stack(_run(net, inputs, pos))          # second call: CUDA-graph replay
    assert a.shape[0] == 1
    assert torch.equal(a, b)
    assert torch.equal(a[0], full[-1])
    # inputs must not be mutated and outputs are fresh tensors by default
    assert all(torch.equal(x, torch.from_numpy(y)) for x, y in zip(inputs, z["inputs"]))


def test_parameter_update_is_picked_up(pvsr_lib):
    """Packed bf16 operands are refreshed when the fp32 master parameters change (e.g. optimizer.step)."""
    z, meta = load_golden("x2_pos")
    net = build_net(meta["kwargs"]).cuda().eval()
    inputs = [torch.from_numpy(x) for x in z["inputs"]]
    pos = torch.from_numpy(z["pos"])
    a = _stack(_run(net, inputs, pos))
    with torch.no_grad():
        net.out_block.conv2.bias.add_(1.0)
    b = _stack(_run(net, inputs, pos))
    assert torch.allclose(b, a + 1.0, atol=1e-5)
    with torch.no_grad():
        net.forward_lstm_block.cell_list[0].conv.weight.mul_(0.5)
    c = _stack(_run(net, inputs, pos))
    assert (c - b).abs().max().item() > 1e-4


def test_errors_like_reference(pvsr_lib):
    from src.model.nets import RefineNet
    with pytest.raises(ValueError):
        RefineNet(1, 1, [64, 64, 64], upscale_factor=5)
    with pytest.raises(ValueError):
        RefineNet(1, 1, [64, 64, 64], update_memory=False, num_updated_frames=6)
    net = RefineNet(1, 1, [64, 64, 64], num_stages=1, update_memory=False, num_updated_frames=0).cuda().eval()
    with pytest.raises(IndexError):
        with torch.no_grad():
            net([torch.zeros(1, 1, 8, 8, device="cuda")] * 5, torch.zeros(1, 5, 1, device="cuda"))
    net2 = build_net(dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=1, update_memory=True,
                          num_updated_frames=2, upscale_factor=2))
    from pvsr.lib import PvsrError
    with pytest.raises(PvsrError):
        with torch.no_grad():
            net2.eval()([torch.zeros(1, 1, 8, 8)] * 6, torch.zeros(1, 6, 1))   # CPU tensors: no fallback


@pytest.mark.parametrize("scale,h,w,pos", [(4, 54, 63, True), (4, 63, 48, True), (3, 72, 84, True),
                                           (2, 108, 126, True)])   # BASELINE.json configs[1..3]: ACDC x4/x3/x2, DSB15 x4
def test_full_size_sequence_vs_oracle(pvsr_lib, scale, h, w, pos):
    """One ACDCSR / DSB15SR-shaped cine sequence (T=30, U=6): SR frames and PSNR/SSIM against the CPU oracle."""
    from oracle import refinenet_oracle as O
    T, U = 30, 6
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=3, update_memory=True,
              num_updated_frames=U, refine_window_size=5, upscale_factor=scale, positional_encoding=pos)
    net = build_net(kw)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    frames = [torch.randn(1, 1, h, w, generator=g) for _ in range(T)]
    inputs = O.circular_window(frames, T, U)
    code = torch.from_numpy(O.positional_code(T, 11))
    pos_codes = torch.stack(O.circular_window(list(code), T, U)).view(1, T + 2 * U, 1)
    hr = [torch.randn(1, 1, h * scale, w * scale, generator=g) for _ in range(T)]
    with torch.no_grad():
        ref = O.refinenet_forward(sd, inputs, pos_codes, **oracle_kwargs(kw))[-1]
    net = net.cuda().eval()
    net.only_last_head = True
    got = _run(net, inputs, pos_codes)[-1]
    got = [o.cpu() for o in got]
    errs = [(a - b).abs().max().item() for a, b in zip(got, ref)]
    rels = [((a - b).norm() / b.norm()).item() for a, b in zip(got, ref)]
    assert max(errs) <= MAX_ABS and max(rels) <= REL_L2, (max(errs), max(rels))
    # metric parity on denormalised frames (reference metrics.py / utils.py semantics)
    dataset = "acdc"
    dp, ds = [], []
    for a, b, t in zip(got, ref, hr):
        ta = O.denormalize(t, dataset)
        dp.append(abs(float(O.psnr(O.denormalize(a, dataset), ta)) - float(O.psnr(O.denormalize(b, dataset), ta))))
        ds.append(abs(float(O.ssim(O.denormalize(a, dataset), ta)) - float(O.ssim(O.denormalize(b, dataset), ta))))
    # SSIM gate = 2x the drift of the REFERENCE ITSELF under torch.autocast(bf16) vs fp32 on the same sequence and
    # target (SURVEY 8c derived 1e-4 that way at x4; measured here with the unmodified reference, CPU, seed 1234:
    # x4 54x63: SSIM 4.9e-5, PSNR 2.2e-4 dB;  x2 108x126: SSIM 1.02e-4, PSNR 3.2e-4 dB - the x2 frames are random-noise
    # targets 4x larger than the LR grid, where single uint8 flips move the 11x11-window SSIM more)
    ssim_gate = 2e-4 if scale == 2 else 1e-4
    assert max(dp) <= 0.01 and max(ds) <= ssim_gate, (max(dp), max(ds))


def test_batched_sequences_are_independent(pvsr_lib):
    """Batching N sequences (the inference sharding unit) gives the same frames as running them one by one."""
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=3, update_memory=True,
              num_updated_frames=3, refine_window_size=5, upscale_factor=4, positional_encoding=True)
    net = build_net(kw).cuda().eval()
    net.only_last_head = True
    g = torch.Generator().manual_seed(5)
    N, L, h, w = 3, 10, 20, 17
    inputs = [torch.randn(N, 1, h, w, generator=g) for _ in range(L)]
    pos = torch.randn(N, L, 1, generator=g)
    full = _stack(_run(net, inputs, pos))[0]
    for n in range(N):
        one = _stack(_run(net, [x[n:n + 1] for x in inputs], pos[n:n + 1]))[0]
        assert torch.allclose(one[:, 0], full[:, n], atol=1e-6), n


def test_host_frame_ring_overlapped_readback(pvsr_lib):
    """pvsr.hostio.HostFrameRing: frames of consecutive steps (engine output buffer reused) arrive intact."""
    from pvsr.hostio import HostFrameRing
    from src.model.nets import RefineNet
    torch.manual_seed(0)
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=2, update_memory=True,
              num_updated_frames=3, refine_window_size=5, upscale_factor=2, positional_encoding=True)
    net = RefineNet(**kw).to("cuda").eval()
    net.only_last_head = True
    net.reuse_output_buffers = True
    ring = HostFrameRing("cuda", slots=2)
    g = torch.Generator().manual_seed(5)
    batches = [([torch.randn(2, 1, 9, 11, generator=g) for _ in range(8)], torch.randn(2, 8, 1, generator=g))
               for _ in range(4)]
    want, slots = [], []
    with torch.no_grad():
        for xs, ps in batches:                       # reference results, one step at a time
            want.append(torch.stack([f.clone() for f in net([x.cuda() for x in xs], ps.cuda())[-1]]).cpu())
        got = []
        for i, (xs, ps) in enumerate(batches):       # pipelined: submit, keep going, read two steps later
            ring.before_launch()
            frames = net([x.cuda() for x in xs], ps.cuda())[-1]
            slots.append(ring.submit(frames))
            if i >= 1:
                got.append(ring.result(slots[i - 1]).clone())
        got.append(ring.result(slots[-1]).clone())
        ring.drain()
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_batched_full_size_graph_replay_through_host_ring(pvsr_lib):
    """The benchmarked combination at B = 8: full ACDCSR x4 shape x CUDA-graph replay x last-head-only x rotating output
    buffers x overlapped HostFrameRing read-back, checked against the CPU oracle for 2 of the 8 sequences of the LAST
    of three pipelined steps (so the frames compared went through a replayed graph and a reused host slot)."""
    from oracle import refinenet_oracle as O
    from pvsr.hostio import HostFrameRing
    from pvsr.synthetic import cine_batch
    T, U, h, w, B = 30, 6, 54, 63, 8
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=3, update_memory=True,
              num_updated_frames=U, refine_window_size=5, upscale_factor=4, positional_encoding=True)
    net = build_net(kw)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.cuda().eval()
    net.only_last_head = True
    net.reuse_output_buffers = True
    net.engine.output_slots = 2
    ring = HostFrameRing("cuda", slots=2)
    steps = [cine_batch(B, T=T, U=U, h=h, w=w, scale=4, seed=900 + i) for i in range(3)]
    slot = None
    with torch.no_grad():
        for inputs, pos in steps:
            pl = net.engine.plan_for(B, len(inputs), h, w, False, torch.device("cuda", torch.cuda.current_device()))
            ring.before_launch(net.engine.next_output_ptr(pl))
            frames = net([x.cuda(non_blocking=True) for x in inputs], pos.cuda(non_blocking=True))[-1]
            slot = ring.submit(frames)
        got = ring.result(slot).clone()          # [T, B, 1, H, W] of the last step
        ring.drain()
        inputs, pos = steps[-1]
        for i in (0, B - 1):
            ref = O.refinenet_forward(sd, [x[i:i + 1] for x in inputs], pos[i:i + 1], **oracle_kwargs(kw))[-1]
            for t, r in enumerate(ref):
                g = got[t, i:i + 1]
                assert (g - r).abs().max().item() <= MAX_ABS and ((g - r).norm() / r.norm()).item() <= REL_L2, (i, t)


def _plain_reference_tree(kw):
    """A plain nn.Module tree with ONLY what INTEGRATION.md section B says the engine needs: the reference's attribute
    names (refine_net.py:21-28), its sub-module attributes (`cell.memory`, `refine_block.positional_encoding`) and its
    parameter names - none of this package's RefineNet code."""
    import torch.nn as nn
    s = kw["upscale_factor"]
    net = nn.Module()
    for k in ("in_channels", "out_channels", "num_features", "num_stages", "refine_window_size", "upscale_factor",
              "update_memory", "num_updated_frames"):
        setattr(net, k, kw[k])
    net.in_block = nn.Module()
    net.in_block.conv = nn.Conv2d(1, 64, 3, padding=1)
    net.in_block.prelu = nn.PReLU(1, 0.2)
    for name in ("forward_lstm_block", "backward_lstm_block"):
        blk = nn.Module()
        cells = []
        for _ in kw["num_features"]:
            c = nn.Module()
            c.conv = nn.Conv2d(128, 256, 3, padding=1)
            c.memory = True
            cells.append(c)
        blk.cell_list = nn.ModuleList(cells)
        setattr(net, name, blk)
    rb = nn.Module()
    rb.positional_encoding = True
    rb.body = nn.Sequential()
    rb.body.add_module("conv1", nn.Conv2d(645, 129, 3, padding=1))
    rb.body.add_module("conv2", nn.Conv2d(129, 64, 3, padding=1))
    rb.prelu = nn.PReLU(1, 0.2)
    net.refine_block = rb
    ob = nn.Module()
    n_ps = {2: 1, 3: 1, 4: 2, 8: 3}[s]
    for q in range(n_ps):
        setattr(ob, f"conv{q + 1}", nn.Conv2d(64, 64 * (9 if s == 3 else 4), 3, padding=1))
    setattr(ob, f"conv{n_ps + 1}", nn.Conv2d(64, 1, 3, padding=1))
    net.out_block = ob
    return net


@pytest.mark.parametrize("scale", [2, 3, 4, 8])
def test_engine_binds_the_reference_module(pvsr_lib, scale):
    """`pvsr.engine.bind_reference_module` on (a) the UNMODIFIED reference RefineNet class
    (staged under oracle/_ref, when present) and (b) a plain nn.Module tree with only the attributes section B lists -
    both against the CPU oracle on the module's own weights, every scale (2 / 2 / 3 / 4 head convolutions)."""
    from oracle import ref_model, refinenet_oracle as O
    from pvsr.engine import bind_reference_module
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=2, update_memory=True,
              num_updated_frames=3, refine_window_size=5, upscale_factor=scale, positional_encoding=True)
    g = torch.Generator().manual_seed(11)
    inputs = [torch.randn(2, 1, 9, 7, generator=g) for _ in range(8)]
    pos = torch.randn(2, 8, 1, generator=g)
    nets = [("plain tree", _plain_reference_tree(kw))]
    if ref_model.available():
        torch.manual_seed(3)
        nets.append(("reference class", ref_model.load().RefineNet(**kw)))
    for what, net in nets:
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        assert len(sd) == 26 - 2 * (3 - {2: 2, 3: 2, 4: 3, 8: 4}[scale])
        with torch.no_grad():
            ref = O.refinenet_forward(sd, inputs, pos, **oracle_kwargs(kw))
        net = net.cuda().eval()
        eng = bind_reference_module(net)
        assert net.num_head_convs == {2: 2, 3: 2, 4: 3, 8: 4}[scale] and net.engine is eng
        with torch.no_grad():
            out = eng.forward([x.cuda() for x in inputs], pos.cuda())
        torch.cuda.synchronize()
        assert len(out) == 6
        got, want = _stack(out), torch.stack([torch.stack(list(o)) for o in ref])
        rel = ((got - want).norm() / want.norm()).item()
        assert (got - want).abs().max().item() <= MAX_ABS and rel <= REL_L2, (what, scale, rel)
