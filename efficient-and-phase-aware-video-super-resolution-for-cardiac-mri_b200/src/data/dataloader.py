"""`Dataloader` of the YAML registry (reference src/data/dataloader.py:6-53): torch's DataLoader whose workers
re-seed numpy from the parent's numpy state, plus an optional `shard=(rank, world)` that gives every data-parallel
rank its own slice of the dataset (the reference is single-GPU and has no equivalent).

`DeviceDataloader` (pvsr/device_loader.py) takes the same keywords and serves the same batches from volumes resident
in HBM: one gather kernel per batch instead of numpy workers."""
import numpy as np
from torch.utils.data import DataLoader
from torch.utils.data.distributed import DistributedSampler

from pvsr.parallel import ShardSampler

from pvsr.device_loader import DeviceDataloader  # noqa: F401  (registry name: `dataloader: {name: DeviceDataloader}`)


def _seed_worker(worker_id):
    np.random.seed((int(np.random.get_state()[1][0]) + worker_id) % (2 ** 32))   # int(): numpy 2 keeps uint32 otherwise


class Dataloader(DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, sampler=None, batch_sampler=None, num_workers=0,
                 collate_fn=None, pin_memory=False, drop_last=False, timeout=0, worker_init_fn=None, shard=None,
                 shard_pad=True):
        if shard is not None and sampler is None and batch_sampler is None:
            rank, world = shard
            if world > 1:
                # training: equal step counts on every rank (one all-reduce per step) -> padded shards;
                # shard_pad=False (validation): every sample exactly once, ranks may differ by one item
                sampler = DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=shuffle,
                                             drop_last=drop_last) if shard_pad else ShardSampler(len(dataset), rank, world)
                shuffle = False
        kwargs = dict(batch_size=batch_size, shuffle=shuffle, sampler=sampler, batch_sampler=batch_sampler,
                      num_workers=num_workers, pin_memory=pin_memory, drop_last=drop_last, timeout=timeout,
                      worker_init_fn=worker_init_fn or _seed_worker)
        if collate_fn is not None:
            kwargs['collate_fn'] = collate_fn
        super().__init__(dataset, **kwargs)
