# Trainer registry: names resolved by `getattr(src.runner.trainers, config.trainer.name)` (reference src/main.py:102).
from .base_trainer import BaseTrainer
from .acdc_vsr_refinenet_trainer import AcdcVSRRefineNetTrainer, Dsb15VSRRefineNetTrainer

from .acdc_sisr_trainer import AcdcSISRTrainer, Dsb15SISRTrainer
from .acdc_vsr_trainer import AcdcVSRTrainer, Dsb15VSRTrainer
from .acdc_sisr_srfb_trainer import AcdcSISRSRFBTrainer, Dsb15SISRSRFBTrainer

__all__ = ['BaseTrainer', 'AcdcVSRRefineNetTrainer', 'Dsb15VSRRefineNetTrainer', 'AcdcSISRTrainer', 'Dsb15SISRTrainer',
           'AcdcVSRTrainer', 'Dsb15VSRTrainer', 'AcdcSISRSRFBTrainer', 'Dsb15SISRSRFBTrainer']
