"""Epoch loop, checkpointing and data movement shared by the trainers
(reference src/runner/trainers/base_trainer.py:5-252).

Additions over the reference (which is single-GPU): under torch.distributed every rank trains on its own shard,
per-epoch logs are summed over ranks, and only rank 0 writes logs and checkpoints.  Checkpoints keep the reference's
dictionary layout {'net', 'optimizer', 'lr_scheduler', 'monitor', 'epoch', 'random_state', 'np_random_seeds'} so
files are interchangeable (load needs weights_only=False on torch >= 2.6: the Monitor object is pickled).
"""
import logging
import random

import numpy as np
import torch
from tqdm import tqdm

from pvsr import parallel


def to_device(batch, device):
    """Recursively moves the tensors of a dict / list / tuple batch to `device`."""
    if isinstance(batch, torch.Tensor):
        return batch.to(device, non_blocking=True)
    if isinstance(batch, dict):
        return {k: to_device(v, device) for k, v in batch.items()}
    if isinstance(batch, (list, tuple)):
        return type(batch)(to_device(v, device) for v in batch)
    return batch


class BaseTrainer:
    def __init__(self, device, train_dataloader, valid_dataloader, net, loss_fns, loss_weights, metric_fns, optimizer,
                 lr_scheduler, logger, monitor, num_epochs):
        if isinstance(lr_scheduler, torch.optim.lr_scheduler.CyclicLR):
            raise NotImplementedError('Do not support torch.optim.lr_scheduler.CyclicLR scheduler yet.')
        self.device = device
        self.train_dataloader, self.valid_dataloader = train_dataloader, valid_dataloader
        self.net = net.to(device)
        self.loss_fns = [fn.to(device) for fn in loss_fns]
        self.loss_weights = torch.tensor(loss_weights, dtype=torch.float, device=device)
        self.metric_fns = [fn.to(device) for fn in metric_fns]
        self.optimizer, self.lr_scheduler = optimizer, lr_scheduler
        self.logger, self.monitor = logger, monitor
        self.num_epochs = num_epochs
        self.epoch = 1
        self.np_random_seeds = None
        self.rank, self.world = parallel.rank_world()

    # ------------------------------------------------------------------ epoch loop
    def train(self):
        if self.np_random_seeds is None:
            self.np_random_seeds = random.sample(range(10000000), k=self.num_epochs)
        while self.epoch <= self.num_epochs:
            np.random.seed(self.np_random_seeds[self.epoch - 1] + self.rank)
            if self.world > 1:
                # the augmentation decisions follow the reference's Python `random` calls (src/data/transforms.py); one
                # process = the reference's stream untouched, several ranks = one stream per rank and epoch
                random.seed(self.np_random_seeds[self.epoch - 1] * 1000003 + self.rank)
            logging.info(f'Epoch {self.epoch}.')
            train_log, train_batch, train_outputs = self._run_epoch('training')
            logging.info(f'Train log: {train_log}.')
            valid_log, valid_batch, valid_outputs = self._run_epoch('validation')
            logging.info(f'Valid log: {valid_log}.')

            if self.lr_scheduler is not None:
                if isinstance(self.lr_scheduler, torch.optim.lr_scheduler.ReduceLROnPlateau):
                    self.lr_scheduler.step(valid_log['Loss'])
                else:
                    self.lr_scheduler.step()

            if self.rank == 0 and self.logger is not None:
                self.logger.write(self.epoch, train_log, train_batch, train_outputs, valid_log, valid_batch,
                                  valid_outputs)
            periodic = self.monitor.is_saved(self.epoch)
            best = self.monitor.is_best(valid_log)       # evaluated on every rank: keeps early stopping in sync
            if self.rank == 0:
                for path, what in ((periodic, 'the'), (best, 'the best')):
                    if path:
                        logging.info(f'Save {what} checkpoint to {path}.')
                        self.save(path)
            if self.monitor.is_early_stopped():
                logging.info('Early stopped.')
                break
            self.epoch += 1
        if self.rank == 0 and self.logger is not None:
            self.logger.close()

    def _progress(self, dataloader, mode):
        return tqdm(dataloader, total=len(dataloader), desc=mode, disable=self.rank != 0)

    def _run_epoch(self, mode):
        raise NotImplementedError

    def _allocate_data(self, batch):
        return to_device(batch, self.device)

    def _init_log(self):
        names = ['Loss'] + [fn.__class__.__name__ for fn in self.loss_fns + self.metric_fns]
        return dict.fromkeys(names, 0)

    # ------------------------------------------------------------------ checkpoints
    def save(self, path):
        torch.save({'net': self.net.state_dict(),
                    'optimizer': self.optimizer.state_dict(),
                    'lr_scheduler': self.lr_scheduler.state_dict() if self.lr_scheduler else None,
                    'monitor': self.monitor,
                    'epoch': self.epoch,
                    'random_state': random.getstate(),
                    'np_random_seeds': self.np_random_seeds}, path)

    def load(self, path):
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        self.net.load_state_dict(ckpt['net'])
        self.optimizer.load_state_dict(ckpt['optimizer'])
        if ckpt['lr_scheduler']:
            self.lr_scheduler.load_state_dict(ckpt['lr_scheduler'])
        self.monitor = ckpt['monitor']
        self.epoch = ckpt['epoch'] + 1
        random.setstate(ckpt['random_state'])
        self.np_random_seeds = ckpt['np_random_seeds']
        if hasattr(self.net, 'engine'):
            self.net.engine.params_changed()
