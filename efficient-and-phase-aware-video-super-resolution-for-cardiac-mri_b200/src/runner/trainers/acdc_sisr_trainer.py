"""Trainer of the single-image nets (reference src/runner/trainers/acdc_sisr_trainer.py:8-66 on top of
base_trainer.py:95-145): batch = {'lr_img', 'hr_img'}, losses = loss_fn(output, target), metrics on the
de-normalised images.

Two execution paths produce the same numbers (as in AcdcVSRRefineNetTrainer):
  * generic - `net(input)` is differentiable (pvsr.edsr_engine autograd bridge): the reference's own sequence
              `zero_grad(); loss.backward(); optimizer.step()` with any torch loss / optimiser;
  * fused   - only loss torch.nn.L1Loss + optimiser pvsr.optim.FusedAdam: forward, L1 and backward as CUDA-graph
              replays without autograd, one all-reduce of the flat gradient (data parallel), one Adam kernel.
"""
import functools

import torch

from pvsr import parallel
from pvsr.optim import FusedAdam
from src.utils import denormalize
from ..scores import _fusable, per_sample_scores
from .base_trainer import BaseTrainer


class AcdcSISRTrainer(BaseTrainer):
    dataset_name = 'acdc'

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._denormalize = functools.partial(denormalize, dataset=self.dataset_name)
        fused_opt = isinstance(self.optimizer, FusedAdam)
        self._fused = (fused_opt and hasattr(self.net, 'engine') and hasattr(self.net.engine, 'loss_and_grads')
                       and len(self.loss_fns) == 1 and type(self.loss_fns[0]) is torch.nn.L1Loss
                       and self.loss_fns[0].reduction == 'mean')
        self._dp = parallel.DataParallelStep(self.net, self.optimizer) if fused_opt else None
        self.log_every = 10

    def _run_epoch(self, mode):
        training = mode == 'training'
        self.net.train(training)
        dataloader = self.train_dataloader if training else self.valid_dataloader
        sampler = getattr(dataloader, 'sampler', None)
        if training and hasattr(sampler, 'set_epoch'):
            sampler.set_epoch(self.epoch)
        trange = self._progress(dataloader, mode)
        log, count = self._init_log(), 0
        names = list(log)
        acc = torch.zeros(len(names), dtype=torch.float64, device=self.device)     # weighted sums, device-resident
        batch, outputs = None, None
        for step, batch in enumerate(trange):
            batch = self._allocate_data(batch)
            inputs, targets = self._get_inputs_targets(batch)
            if training and self._fused:
                loss, outputs = self._fused_step(inputs, targets)
                losses = [loss]
                loss = loss * self.loss_weights[0]
                if float(self.loss_weights[0]) != 1.0:
                    self._dp.flat_grad.mul_(self.loss_weights[0])
                self._dp.step()
            elif training:
                outputs = self.net(inputs)
                losses = self._compute_losses(outputs, targets)
                loss = (torch.stack(losses) * self.loss_weights).sum()
                self.optimizer.zero_grad()
                loss.backward()
                self._optimizer_step()
            else:
                with torch.no_grad():
                    outputs = self.net(inputs)
                    losses = self._compute_losses(outputs, targets)
                    loss = (torch.stack(losses) * self.loss_weights).sum()
            metrics = self._compute_metrics(self._detached(outputs), targets)
            batch_size = dataloader.batch_size
            # weighted sums stay on the device; they cross PCIe every `log_every` steps (the reference's per-step
            # .item() calls serialise host and GPU)
            vals = torch.stack([loss.detach().float()] + [l.detach().float() for l in losses] +
                               [m.detach().float() for m in metrics])
            acc += vals.double() * batch_size
            count += batch_size
            if (step + 1) % self.log_every == 0:
                trange.set_postfix(**{k: f'{v / count: .3f}' for k, v in zip(names, acc.tolist())})
        log = dict(zip(names, acc.tolist()))
        log, count = parallel.reduce_log(log, count, self.device)
        return {k: v / max(count, 1) for k, v in log.items()}, batch, outputs

    def _fused_step(self, inputs, targets):
        return self.net.engine.loss_and_grads(inputs, targets)

    @staticmethod
    def _detached(outputs):
        return outputs.detach()

    def _optimizer_step(self):
        if self._dp is not None:
            self._dp.step()
        else:
            if self.world > 1:
                for p in self.net.parameters():
                    if p.grad is not None:
                        parallel.allreduce_sum_(p.grad).mul_(1.0 / self.world)
            self.optimizer.step()

    def _get_inputs_targets(self, batch):
        return batch['lr_img'], batch['hr_img']

    def _compute_losses(self, output, target):
        return [loss_fn(output, target) for loss_fn in self.loss_fns]

    def _compute_metrics(self, output, target):
        with torch.no_grad():
            if _fusable([], self.metric_fns, output) and not any(hasattr(fn, 'inner') for fn in self.metric_fns):
                _, scores = per_sample_scores([], self.metric_fns, output.detach(), target, None, None, None,
                                              dataset=self.dataset_name)
                return list(scores.mean(dim=0))
            output, target = self._denormalize(output), self._denormalize(target)
            return [metric_fn(output, target) for metric_fn in self.metric_fns]


class Dsb15SISRTrainer(AcdcSISRTrainer):
    """Same loop with the DSB15 de-normalisation constants (src/utils.py:15-16)."""
    dataset_name = 'dsb15'
