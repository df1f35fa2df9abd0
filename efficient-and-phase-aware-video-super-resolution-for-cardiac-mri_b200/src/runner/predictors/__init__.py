# Predictor registry: names resolved by `getattr(src.runner.predictors, config.predictor.name)` (src/main.py:152).
from .base_predictor import BasePredictor
from .acdc_vsr_refinenet_predictor import AcdcVSRRefineNetPredictor, Dsb15VSRRefineNetPredictor

from .acdc_sisr_predictor import (AcdcSISRPredictor, AcdcSISRSRFBPredictor, Dsb15SISRPredictor,
                                  Dsb15SISRSRFBPredictor)
from .acdc_vsr_predictor import AcdcVSRPredictor, Dsb15VSRPredictor

__all__ = ['BasePredictor', 'AcdcVSRRefineNetPredictor', 'Dsb15VSRRefineNetPredictor', 'AcdcSISRPredictor',
           'Dsb15SISRPredictor', 'AcdcVSRPredictor', 'Dsb15VSRPredictor',
           'AcdcSISRSRFBPredictor', 'Dsb15SISRSRFBPredictor']
