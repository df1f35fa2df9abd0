"""Trainer of the iterated single-image nets - DRFSISRNet on this path (reference
src/runner/trainers/acdc_sisr_srfb_trainer.py:6-41, dsb15_sisr_srfb_trainer.py): `outputs = net(input)` is a list of
`num_steps` images; every loss is the mean over the steps of `loss_fn(output_s, target)` (:22-26), metrics are taken
on the last output (:39-40).  Loop, logging and the data-parallel step are AcdcSISRTrainer's."""
import torch

from .acdc_sisr_trainer import AcdcSISRTrainer


class AcdcSISRSRFBTrainer(AcdcSISRTrainer):
    dataset_name = 'acdc'

    def _fused_step(self, inputs, targets):
        steps = self.net.num_steps
        return self.net.engine.loss_and_grads([inputs] * steps, [targets] * steps)

    @staticmethod
    def _detached(outputs):
        return [o.detach() for o in outputs]

    def _compute_losses(self, outputs, target):
        return [torch.stack([loss_fn(o, target) for o in outputs]).mean() for loss_fn in self.loss_fns]

    def _compute_metrics(self, outputs, target):
        return super()._compute_metrics(outputs[-1], target)


class Dsb15SISRSRFBTrainer(AcdcSISRSRFBTrainer):
    """Same loop with the DSB15 de-normalisation constants (src/utils.py:15-16)."""
    dataset_name = 'dsb15'
