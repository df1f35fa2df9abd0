"""Data-parallel training over NCCL on 2 GPUs of one box: averaged per-rank gradients equal the single-GPU
gradients of the concatenated batch, ranks end a step with identical weights, and sequence-sharded inference
returns the same SR frames as one GPU.  Skipped when fewer than 2 GPUs are visible (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
KW = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=3, update_memory=True,
          num_updated_frames=3, refine_window_size=5, upscale_factor=4, positional_encoding=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(n):
    import sys
    sys.path.insert(0, PKG)
    from pvsr.synthetic import cine_batch
    return cine_batch(n, T=3, U=3, h=16, w=12, scale=4, seed=11, end_systole=1, with_targets=True)


def _worker(rank, world, port, q):
    import sys
    for p in (PKG, ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    from helpers import build_net
    from pvsr import parallel
    from pvsr.optim import FusedAdam
    parallel.init()
    dev = torch.device("cuda", rank)
    inputs, pos, targets = _batch(2 * world)
    sl = slice(2 * rank, 2 * rank + 2)
    net = build_net(KW, seed=rank).to(dev).train()          # different seeds: the broadcast must equalise them
    opt = FusedAdam.for_net(net, lr=1e-3)
    dp = parallel.DataParallelStep(net, opt)
    loss, _ = net.engine.loss_and_grads([x[sl].to(dev) for x in inputs], pos[sl].to(dev),
                                        [t[sl].to(dev) for t in targets])
    parallel.allreduce_sum_(dp.flat_grad)
    grads = {k: (p.grad / world).cpu() for k, p in net.named_parameters()}
    opt.step()
    net.engine.params_changed()
    torch.cuda.synchronize()
    weights = {k: p.detach().cpu() for k, p in net.named_parameters()}
    # sequence-sharded inference: this rank's share of 4 sequences
    net.eval()
    net.only_last_head = True
    mine = parallel.shard_indices(2 * world, rank, world)
    with torch.no_grad():
        out = net([x[mine].to(dev) for x in inputs], pos[mine].to(dev))[-1]
    frames = torch.stack([o.cpu() for o in out])
    q.put((rank, loss.item(), grads, weights, mine, frames))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_data_parallel_matches_single_gpu(pvsr_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from helpers import build_net
    from pvsr.optim import FusedAdam
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
    assert all(p.exitcode == 0 for p in procs)

    inputs, pos, targets = _batch(2 * world)
    net = build_net(KW, seed=0).cuda().train()
    opt = FusedAdam.for_net(net, lr=1e-3)
    loss, _ = net.engine.loss_and_grads([x.cuda() for x in inputs], pos.cuda(), [t.cuda() for t in targets])
    torch.cuda.synchronize()
    assert abs(sum(r[1] for r in res) / world - loss.item()) <= 1e-4 * abs(loss.item())
    for k, p in net.named_parameters():
        g = p.grad.cpu()
        for r in res:
            gr = r[2][k]
            if float(g.abs().sum()) == 0.0:
                assert float(gr.abs().sum()) == 0.0
                continue
            rel = ((gr - g).norm() / g.norm()).item()
            assert rel <= 2e-3, (k, rel)            # same kernels; batch split changes the fp32 summation order only
        assert torch.equal(res[0][2][k], res[1][2][k])          # all-reduce leaves identical gradients on every rank
    opt.step()
    torch.cuda.synchronize()
    for k, p in net.named_parameters():
        assert torch.equal(res[0][3][k], res[1][3][k]), k        # ranks stay in lock-step
    # sharded inference == single-GPU inference, sequence by sequence
    net.engine.params_changed()
    net.eval()
    net.only_last_head = True
    with torch.no_grad():
        full = torch.stack([o.cpu() for o in net([x.cuda() for x in inputs], pos.cuda())[-1]])
    for r in res:
        assert r[4] == list(range(r[0], 2 * world, world))
        # rank weights after the step equal the single-GPU weights up to Adam's sensitivity; compare loosely
        assert ((r[5] - full[:, r[4]]).norm() / full[:, r[4]].norm()).item() <= 2e-2


# ------------------------------------------------------------------------------------------------ EDSR
EDSR_KW = dict(in_channels=1, out_channels=1, num_resblocks=2, num_features=64, upscale_factor=4, res_scale=0.1)


def _edsr_batch(n):
    g = torch.Generator().manual_seed(21)
    return torch.randn(n, 1, 12, 10, generator=g), torch.randn(n, 1, 48, 40, generator=g)


def _edsr_worker(rank, world, port, q):
    import sys
    for p in (PKG, ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    from pvsr import parallel
    from pvsr.optim import FusedAdam
    from src.model.nets import EDSRNet
    parallel.init()
    dev = torch.device("cuda", rank)
    x, t = _edsr_batch(2 * world)
    sl = slice(2 * rank, 2 * rank + 2)
    torch.manual_seed(rank)                                  # different seeds: the broadcast must equalise them
    net = EDSRNet(**EDSR_KW).to(dev).train()
    opt = FusedAdam.for_net(net, lr=1e-3)
    dp = parallel.DataParallelStep(net, opt)
    loss, _ = net.engine.loss_and_grads(x[sl].to(dev), t[sl].to(dev))
    parallel.allreduce_sum_(dp.flat_grad)
    grads = {k: (p.grad / world).cpu() for k, p in net.named_parameters()}
    opt.step()
    net.engine.params_changed()
    torch.cuda.synchronize()
    weights = {k: p.detach().cpu() for k, p in net.named_parameters()}
    q.put((rank, loss.item(), grads, weights))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_edsr_data_parallel_matches_single_gpu(pvsr_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pvsr.optim import FusedAdam
    from src.model.nets import EDSRNet
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_edsr_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    x, t = _edsr_batch(2 * world)
    torch.manual_seed(0)
    net = EDSRNet(**EDSR_KW).cuda().train()
    FusedAdam.for_net(net, lr=1e-3)
    loss, _ = net.engine.loss_and_grads(x.cuda(), t.cuda())
    torch.cuda.synchronize()
    # nn.L1Loss averages over the batch: the mean of the two half-batch losses is the full-batch loss
    assert abs(sum(r[1] for r in res) / world - loss.item()) <= 1e-4 * abs(loss.item())
    for k, p in net.named_parameters():
        g = p.grad.cpu()
        for r in res:
            rel = ((r[2][k] - g).norm() / g.norm()).item()
            assert rel <= 5e-3, (k, rel)
        assert torch.equal(res[0][2][k], res[1][2][k])
        assert torch.equal(res[0][3][k], res[1][3][k]), k        # ranks stay in lock-step


# ------------------------------------------------------------------------------------------------ DRFNet
DRF_KW = dict(in_channels=1, out_channels=1, num_features=64, num_groups=2, upscale_factor=4)


def _drf_batch(n, T=3):
    g = torch.Generator().manual_seed(23)
    return ([torch.randn(n, 1, 12, 10, generator=g) for _ in range(T)],
            [torch.randn(n, 1, 48, 40, generator=g) for _ in range(T)])


def _drf_worker(rank, world, port, q):
    import sys
    for p in (PKG, ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    from pvsr import parallel
    from pvsr.optim import FusedAdam
    from src.model.nets import DRFNet
    parallel.init()
    dev = torch.device("cuda", rank)
    xs, ts = _drf_batch(2 * world)
    sl = slice(2 * rank, 2 * rank + 2)
    torch.manual_seed(rank)                                  # different seeds: the broadcast must equalise them
    net = DRFNet(**DRF_KW).to(dev).train()
    opt = FusedAdam.for_net(net, lr=1e-3)
    dp = parallel.DataParallelStep(net, opt)
    loss, _ = net.engine.loss_and_grads([x[sl].to(dev) for x in xs], [t[sl].to(dev) for t in ts])
    parallel.allreduce_sum_(dp.flat_grad)
    grads = {k: (p.grad / world).cpu() for k, p in net.named_parameters()}
    opt.step()
    net.engine.params_changed()
    torch.cuda.synchronize()
    weights = {k: p.detach().cpu() for k, p in net.named_parameters()}
    q.put((rank, loss.item(), grads, weights))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_drfnet_data_parallel_matches_single_gpu(pvsr_lib):
    """Averaged BPTT gradients of two ranks (two sequences each) equal the single-GPU gradients of the four sequences;
    the ranks stay bit-identical after the fused Adam step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pvsr.optim import FusedAdam
    from src.model.nets import DRFNet
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_drf_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    xs, ts = _drf_batch(2 * world)
    torch.manual_seed(0)
    net = DRFNet(**DRF_KW).cuda().train()
    FusedAdam.for_net(net, lr=1e-3)
    loss, _ = net.engine.loss_and_grads([x.cuda() for x in xs], [t.cuda() for t in ts])
    torch.cuda.synchronize()
    assert abs(sum(r[1] for r in res) / world - loss.item()) <= 1e-4 * abs(loss.item())
    for k, p in net.named_parameters():
        g = p.grad.cpu()
        if p.numel() > 1:                                    # slope gradients are cancellation-dominated scalars
            for r in res:
                rel = ((r[2][k] - g).norm() / g.norm()).item()
                assert rel <= 2e-2, (k, rel)
        assert torch.equal(res[0][2][k], res[1][2][k])
        assert torch.equal(res[0][3][k], res[1][3][k]), k        # ranks stay in lock-step
