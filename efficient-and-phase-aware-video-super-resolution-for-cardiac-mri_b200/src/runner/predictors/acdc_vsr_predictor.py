"""Test loop of the plain video-SR nets - DRFNet on this path (reference
src/runner/predictors/acdc_vsr_predictor.py:15-179, dsb15_vsr_predictor.py): `outputs = net(inputs)` (:62), per-frame
losses / metrics / exports exactly as the RefineNet predictor, whose loop (shape-bucketed batching of sequences,
sequence sharding over ranks, one device->host transfer of all scalars per launch) this class reuses."""
from .acdc_vsr_refinenet_predictor import AcdcVSRRefineNetPredictor


class AcdcVSRPredictor(AcdcVSRRefineNetPredictor):
    dataset_name = 'acdc'

    def _forward(self, inputs, pos_codes):
        return self.net(inputs)

    def _get_inputs_targets(self, batch):
        return batch['lr_imgs'], batch['hr_imgs'], None, batch['index']


class Dsb15VSRPredictor(AcdcVSRPredictor):
    """Same loop with the DSB15 de-normalisation constants (src/utils.py:15-16)."""
    dataset_name = 'dsb15'
