"""Trainer of the plain video-SR nets - DRFNet on this path (reference src/runner/trainers/acdc_vsr_trainer.py:9-122,
dsb15_vsr_trainer.py).

`outputs = net(inputs)` is ONE list of T frames; every loss is the mean over the T frames of `loss_fn(output_t,
target_t)` (:83-87), metrics are PSNR / SSIM on the de-normalised frames averaged over frames (:99-107).  The loop,
the logging and the data-parallel step are AcdcVSRRefineNetTrainer's; as there, two execution paths give the same
numbers: generic autograd (`net(inputs)` is differentiable, any torch loss / optimiser) and the fused path (one
torch.nn.L1Loss + pvsr.optim.FusedAdam: forward, L1 and BPTT backward without autograd, one all-reduce, one Adam kernel).
"""
import torch

from .acdc_vsr_refinenet_trainer import AcdcVSRRefineNetTrainer


class AcdcVSRTrainer(AcdcVSRRefineNetTrainer):
    dataset_name = 'acdc'

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._fused = self._fused and hasattr(getattr(self.net, 'engine', None), 'loss_and_grads')

    def _get_inputs_targets(self, batch):
        return batch['lr_imgs'], batch['hr_imgs'], None

    def _forward(self, inputs, pos_codes):
        return (self.net(inputs),)             # a 1-tuple of lists: the shared loop reads outputs[-1]

    def _fused_step(self, inputs, targets, pos_codes):
        loss, outs = self.net.engine.loss_and_grads(inputs, targets)
        loss = loss * self.loss_weights[0]
        if float(self.loss_weights[0]) != 1.0:
            self._dp.flat_grad.mul_(self.loss_weights[0])
        self._dp.step()
        return (outs,), loss, [loss / self.loss_weights[0]]

    def _compute_losses(self, outputs, targets):
        return [torch.stack([loss_fn(o, t) for o, t in zip(outputs[-1], targets)]).mean() for loss_fn in self.loss_fns]


class Dsb15VSRTrainer(AcdcVSRTrainer):
    """Same loop with the DSB15 de-normalisation constants (src/utils.py:15-16)."""
    dataset_name = 'dsb15'
