"""Multi-GPU plumbing of the RefineNet path: one process per GPU (torchrun), torch.distributed over NCCL on the
B200 box (gloo in the CPU tests).

  * inference: cine sequences are independent units (reference predictor loop,
    src/runner/predictors/acdc_vsr_refinenet_predictor.py:53-62) -> sharded by sequence, no data-path collective;
    only the scalar metric sums are all-reduced;
  * training : plain data parallelism (no cross-sample statistics anywhere in refine_net.py) -> ONE all-reduce over
    the flat fp32 gradient buffer per step (weights are shared over stages and time steps, so no gradient is final
    before the end of backward: there is nothing to overlap), averaging folded into FusedAdam's grad_scale.

The reference has no distributed code at all; this module is the added scale-out layer.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')),
            int(os.environ.get('WORLD_SIZE', '1')))


def init(backend=None, device=None):
    """Initialises the default process group when WORLD_SIZE > 1.  Returns (rank, world)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            if device is None:
                device = torch.device('cuda', local_rank)
            torch.cuda.set_device(device)
            kw['device_id'] = device
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_world():
    return (dist.get_rank(), dist.get_world_size()) if is_distributed() else (0, 1)


def shard_indices(n_items, rank, world, sizes=None):
    """Indices of the items (cine sequences) rank `rank` processes.

    Without `sizes`: round-robin (item i -> rank i % world), which keeps per-rank counts within one of each other.
    With `sizes` (a per-item cost, e.g. frames x pixels): greedy longest-processing-time assignment, so ranks finish
    together when sequences differ in length or resolution; ties broken by index for determinism."""
    if world <= 1:
        return list(range(n_items))
    if sizes is None:
        return list(range(rank, n_items, world))
    order = sorted(range(n_items), key=lambda i: (-sizes[i], i))
    load = [0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += sizes[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


class ShardSampler(torch.utils.data.Sampler):
    """Rank `rank`'s items of a dataset WITHOUT padding: i -> rank i % world, in order.  For loops with no per-step
    collective (validation, test): DistributedSampler would repeat samples to equalise the shards, and the summed
    log would then count them twice (the validation score feeds Monitor.is_best and ReduceLROnPlateau)."""

    def __init__(self, n_items, rank, world):
        self.indices = shard_indices(n_items, rank, world)

    def __iter__(self):
        return iter(self.indices)

    def __len__(self):
        return len(self.indices)


def bucket_by_shape(shapes):
    """Groups item indices by identical (frames, h, w): sequences of one bucket can be batched into one plan launch
    (the plan geometry is static).  Returns {shape: [indices]} with deterministic ordering."""
    buckets = {}
    for i, s in enumerate(shapes):
        buckets.setdefault(tuple(s), []).append(i)
    return buckets


def allreduce_sum_(tensor):
    """In-place SUM all-reduce (no-op for a single process)."""
    if is_distributed():
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def broadcast_(tensor, src=0):
    if is_distributed():
        dist.broadcast(tensor, src=src)
    return tensor


def reduce_log(log, count, device):
    """Sums a {name: weighted sum} log dict and its sample count over all ranks (metric bookkeeping of the
    runners); returns (log, count) as Python numbers."""
    if not is_distributed():
        return log, count
    keys = sorted(log)
    t = torch.tensor([float(log[k]) for k in keys] + [float(count)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    vals = t.tolist()
    return {k: v for k, v in zip(keys, vals[:-1])}, vals[-1]


class DataParallelStep:
    """Gradient exchange + optimiser step of data-parallel RefineNet training.

    `flat_grad` is the single fp32 buffer every parameter's .grad is a view of (engine.flatten_parameters()).
    step(): all-reduce(SUM) it over NCCL / NVLink, then FusedAdam with grad_scale = 1 / world (the average)."""

    def __init__(self, net, optimizer):
        self.net, self.optimizer = net, optimizer
        self.flat_param, self.flat_grad = net.engine.flatten_parameters()
        _, self.world = rank_world()
        if hasattr(optimizer, 'grad_scale'):
            optimizer.grad_scale = 1.0 / self.world
        broadcast_(self.flat_param, 0)      # identical initial weights on every rank
        net.engine.params_changed()
        # optional device timing of the exchange: set `time_allreduce = True`, read allreduce_ms() after a synchronize
        self.time_allreduce = False
        self._ar_events = []

    def allreduce_bytes(self):
        return self.flat_grad.numel() * self.flat_grad.element_size()

    def allreduce_ms(self, reset=True):
        """Mean device time (CUDA events on the compute stream, which waits for NCCL's stream) of the gradient
        all-reduce over the steps recorded since the last reset; None when nothing was recorded."""
        if not self._ar_events:
            return None
        ms = [a.elapsed_time(b) for a, b in self._ar_events]
        if reset:
            self._ar_events = []
        return sum(ms) / len(ms)

    def step(self):
        if self.time_allreduce and is_distributed():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            allreduce_sum_(self.flat_grad)
            e1.record()
            self._ar_events.append((e0, e1))
        else:
            allreduce_sum_(self.flat_grad)
        if not hasattr(self.optimizer, 'grad_scale') and self.world > 1:
            self.flat_grad.mul_(1.0 / self.world)
        self.optimizer.step()
        self.net.engine.params_changed()
