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
