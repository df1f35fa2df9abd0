from .base_logger import BaseLogger
from .acdc_vsr_logger import AcdcVSRLogger

from .acdc_sisr_logger import AcdcSISRLogger, AcdcSISRSRFBLogger, Dsb15SISRLogger, Dsb15SISRSRFBLogger

__all__ = ['BaseLogger', 'AcdcVSRLogger', 'AcdcSISRLogger', 'Dsb15SISRLogger', 'AcdcSISRSRFBLogger', 'Dsb15SISRSRFBLogger']
