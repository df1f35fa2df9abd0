"""Trainer of RefineNet (reference src/runner/trainers/acdc_vsr_refinenet_trainer.py:10-136).

Loss (training): for each of the 3*S output lists k, the per-frame losses are averaged over the T frames and weighted
by 0.5 ** (S - 1 - k // 3); the lists are summed (:83-93).  Validation: the last list only (:95-100).
Metrics: PSNR / SSIM on the de-normalised last list, averaged over frames (:103-120).

Two execution paths produce the same numbers:
  * generic  - `net(inputs, pos_codes)` returns differentiable tensors (pvsr.autograd), any torch loss / optimiser
               works, exactly the reference's sequence `zero_grad(); loss.backward(); optimizer.step()`;
  * fused    - taken automatically when the only loss is torch.nn.L1Loss and the optimiser is pvsr.optim.FusedAdam:
               forward, multi-stage L1 and backward run as two CUDA-graph replays without autograd, the gradient
               all-reduce (data parallel) and Adam follow as one collective and one kernel.
"""
import functools

import numpy as np
import torch

from pvsr import parallel
from pvsr.optim import FusedAdam
from src.utils import denormalize
from ..scores import _fusable, per_sample_scores
from .base_trainer import BaseTrainer


class AcdcVSRRefineNetTrainer(BaseTrainer):
    dataset_name = 'acdc'

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._denormalize = functools.partial(denormalize, dataset=self.dataset_name)
        self._fused = (isinstance(self.optimizer, FusedAdam) and len(self.loss_fns) == 1
                       and type(self.loss_fns[0]) is torch.nn.L1Loss and self.loss_fns[0].reduction == 'mean')
        self._dp = parallel.DataParallelStep(self.net, self.optimizer) if isinstance(self.optimizer, FusedAdam) else None
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
            inputs, targets, pos_codes = self._get_inputs_targets(batch)
            T = len(inputs)
            if training and self._fused:
                outputs, loss, losses = self._fused_step(inputs, targets, pos_codes)
            elif training:
                outputs = self._forward(inputs, pos_codes)
                losses = self._compute_losses(outputs, targets)
                loss = (torch.stack(losses) * self.loss_weights).sum()
                self.optimizer.zero_grad()
                loss.backward()
                self._optimizer_step()
            else:
                with torch.no_grad():
                    outputs = self._forward(inputs, pos_codes)
                    losses = self._compute_losses(outputs, targets)
                    loss = (torch.stack(losses) * self.loss_weights).sum()
            metrics = self._compute_metrics(outputs, targets)
            batch_size = dataloader.batch_size
            # the reference reads every scalar back with .item() each step (:60-66), which serialises host and GPU;
            # here the weighted sums stay on the device and cross PCIe every `log_every` steps for the progress bar
            vals = torch.stack([loss.detach().float()] + [l.detach().float() for l in losses] +
                               [m.detach().float() for m in metrics])
            acc += vals.double() * (batch_size * T)
            count += batch_size * T
            if (step + 1) % self.log_every == 0:
                trange.set_postfix(**{k: f'{v / count: .3f}' for k, v in zip(names, acc.tolist())})
        log = dict(zip(names, acc.tolist()))
        log, count = parallel.reduce_log(log, count, self.device)
        return {k: v / max(count, 1) for k, v in log.items()}, batch, outputs[-1] if outputs is not None else None

    def _forward(self, inputs, pos_codes):
        """The tuple of output lists of the net (:44,51)."""
        return self.net(inputs, pos_codes)

    def _optimizer_step(self):
        if self._dp is not None:
            self._dp.step()
        else:
            if self.world > 1:                        # torch optimiser under data parallelism: average the grads
                for p in self.net.parameters():
                    if p.grad is not None:
                        parallel.allreduce_sum_(p.grad).mul_(1.0 / self.world)
            self.optimizer.step()

    def _fused_step(self, inputs, targets, pos_codes):
        loss, out = self.net.engine.loss_and_grads(inputs, pos_codes, targets)
        loss = loss * self.loss_weights[0]
        if float(self.loss_weights[0]) != 1.0:
            self._dp.flat_grad.mul_(self.loss_weights[0])
        self._dp.step()
        outputs = tuple([out[l, t].unsqueeze(1) for t in range(out.shape[1])] for l in range(out.shape[0]))
        return outputs, loss, [loss / self.loss_weights[0]]

    def _get_inputs_targets(self, batch):
        return batch['lr_imgs'], batch['hr_imgs'], batch['pos_code']

    def _compute_losses(self, outputs, targets):
        losses = []
        if self.net.training:
            n_stages = len(outputs) // 3
            for loss_fn in self.loss_fns:
                per_list = []
                for k, frames in enumerate(outputs):
                    discount = np.power(0.5, n_stages - k // 3 - 1)
                    per_list.append(torch.stack([loss_fn(o, t) * discount for o, t in zip(frames, targets)]).mean())
                losses.append(torch.stack(per_list).sum())
        else:
            for loss_fn in self.loss_fns:
                losses.append(torch.stack([loss_fn(o, t) for o, t in zip(outputs[-1], targets)]).mean())
        return losses

    def _compute_metrics(self, outputs, targets):
        """Mean over the T frames of each metric's batch mean (:103-120) - all T * N frames in ONE call per metric
        (equal sample counts per frame make the two means identical)."""
        with torch.no_grad():
            out = torch.stack([o.detach() for o in outputs[-1]]).flatten(0, 1)
            tgt = torch.stack(list(targets)).flatten(0, 1)
            if _fusable([], self.metric_fns, out) and not any(hasattr(fn, 'inner') for fn in self.metric_fns):
                # PSNR / SSIM of all frames in two launches (pvsr_frame_scores)
                _, scores = per_sample_scores([], self.metric_fns, out, tgt, None, None, None,
                                              dataset=self.dataset_name)
                return list(scores.mean(dim=0))
            sr, hr = self._denormalize(out), self._denormalize(tgt)
            return [fn(sr, hr).mean() for fn in self.metric_fns]

    def _update_log(self, log, batch_size, T, loss, losses, metrics):
        weight = batch_size * T
        # one device->host transfer for all scalars of the step (the reference does one .item() per entry)
        vals = torch.stack([loss.detach().float()] + [l.detach().float() for l in losses] +
                           [m.detach().float() for m in metrics]).tolist()
        names = ['Loss'] + [fn.__class__.__name__ for fn in self.loss_fns + self.metric_fns]
        for name, v in zip(names, vals):
            log[name] += v * weight


class Dsb15VSRRefineNetTrainer(AcdcVSRRefineNetTrainer):
    """Same loop with the DSB15 de-normalisation constants (src/utils.py:15-16)."""
    dataset_name = 'dsb15'
