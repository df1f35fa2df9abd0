from .base_logger import BaseLogger
from .acdc_vsr_logger import AcdcVSRLogger

__all__ = ['BaseLogger', 'AcdcVSRLogger']
