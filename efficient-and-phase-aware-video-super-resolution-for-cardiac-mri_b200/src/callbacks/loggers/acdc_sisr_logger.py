"""HR | SR panels of the last batch (reference src/callbacks/loggers/acdc_sisr_logger.py:7-32)."""
import torch

from .acdc_vsr_logger import _column
from .base_logger import BaseLogger


class AcdcSISRLogger(BaseLogger):
    def _add_images(self, epoch, train_batch, train_output, valid_batch, valid_output):
        for tag, batch, output in (('train', train_batch, train_output), ('valid', valid_batch, valid_output)):
            self._add_image(tag, torch.cat([_column(batch['hr_img']), _column(output)], dim=-1), epoch)


class Dsb15SISRLogger(AcdcSISRLogger):
    pass


class AcdcSISRSRFBLogger(AcdcSISRLogger):
    """Iterated nets (reference acdc_sisr_srfb_logger.py:13-31): the panel shows the last of the `num_steps` outputs."""

    def _add_images(self, epoch, train_batch, train_outputs, valid_batch, valid_outputs):
        super()._add_images(epoch, train_batch, train_outputs[-1], valid_batch, valid_outputs[-1])


class Dsb15SISRSRFBLogger(AcdcSISRSRFBLogger):
    pass
