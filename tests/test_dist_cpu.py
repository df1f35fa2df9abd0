"""world_size-2 gloo tests (CPU) of the multi-process host logic: sequence sharding covers the job exactly once,
metric logs and the flat gradient buffer are summed over ranks, initial weights are broadcast from rank 0, CSV rows
are gathered on rank 0.  The NCCL path runs the same code on the GPU box (tests/test_train_gpu.py, bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    from pvsr import parallel
    r, w = parallel.init(backend="gloo")
    assert (r, w) == (rank, world) and parallel.is_distributed() and parallel.rank_world() == (rank, world)
    # 1. sequence sharding: disjoint cover
    mine = parallel.shard_indices(11, r, w)
    gathered = [None] * w
    dist.all_gather_object(gathered, mine)
    assert sorted(i for p in gathered for i in p) == list(range(11))
    # 2. metric log reduction (weighted sums + counts)
    log, count = parallel.reduce_log({'Loss': 1.0 + r, 'PSNR': 30.0 * (r + 1)}, 10 * (r + 1), torch.device('cpu'))
    assert log == {'Loss': 3.0, 'PSNR': 90.0} and count == 30.0
    # 3. flat gradient exchange + broadcast of the initial weights
    flat_g = torch.full((1024,), float(r + 1))
    parallel.allreduce_sum_(flat_g)
    assert torch.all(flat_g == 3.0)
    flat_p = torch.full((1024,), float(r + 7))
    parallel.broadcast_(flat_p, 0)
    assert torch.all(flat_p == 7.0)
    # 4. CSV rows gathered on rank 0
    rows = [[f'seq{r}_frame{t:02d}', 30.0 + t] for t in range(2)]
    out = [None] * w if r == 0 else None
    dist.gather_object(rows, out, dst=0)
    if r == 0:
        assert sorted(x[0] for part in out for x in part) == ['seq0_frame00', 'seq0_frame01', 'seq1_frame00', 'seq1_frame01']
    # 5. data-parallel averaging equals the gradient of the concatenated batch for a mean-reduced loss
    torch.manual_seed(0)
    wgt = torch.randn(8, requires_grad=True)
    data = torch.arange(32, dtype=torch.float32).view(4, 8) / 10
    ((data[2 * r:2 * r + 2] * wgt).sum(dim=1) ** 2).mean().backward()
    g = wgt.grad.clone()
    parallel.allreduce_sum_(g).mul_(1.0 / w)
    wgt.grad = None
    ((data * wgt).sum(dim=1) ** 2).mean().backward()
    assert torch.allclose(g, wgt.grad, atol=1e-5)
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


@pytest.mark.timeout(120)
def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]


def test_unpadded_validation_shards_and_slice_level_sisr_shards():
    """ShardSampler: every item exactly once over the ranks (no DistributedSampler padding in validation / test);
    AcdcSISRPredictor._my_indices: whole (patient, slice) groups per rank, so one rank writes each slice's GIF."""
    import types
    from pvsr.parallel import ShardSampler
    from src.runner.predictors.acdc_sisr_predictor import AcdcSISRPredictor
    for n, world in ((7, 2), (10, 4), (3, 8)):
        seen = [i for r in range(world) for i in ShardSampler(n, r, world)]
        assert sorted(seen) == list(range(n))
    # 3 patients x 2 slices x ragged frame counts
    names = []
    for p in range(3):
        for s in range(2):
            for f in range(4 + p + s):
                names.append((f'patient{p:03d}_2d_slice{s:02d}_frame{f:02d}', f'patient{p:03d}', f'slice{s:02d}', f'frame{f:02d}'))
    owners = {}
    covered = []
    for rank in range(4):
        stub = types.SimpleNamespace(world=4, rank=rank, _name=lambda i: names[i])
        mine = AcdcSISRPredictor._my_indices(stub, len(names))
        covered.extend(mine)
        for i in mine:
            assert owners.setdefault(names[i][1:3], rank) == rank          # a slice never spans two ranks
    assert sorted(covered) == list(range(len(names)))
    one = AcdcSISRPredictor._my_indices(types.SimpleNamespace(world=1, rank=0, _name=lambda i: names[i]), len(names))
    assert one == list(range(len(names)))
