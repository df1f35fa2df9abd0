from .base_logger import BaseLogger
from .acdc_vsr_logger import AcdcVSRLogger

from .acdc_sisr_logger import AcdcSISRLogger, Dsb15SISRLogger

__all__ = ['BaseLogger', 'AcdcVSRLogger', 'AcdcSISRLogger', 'Dsb15SISRLogger']
