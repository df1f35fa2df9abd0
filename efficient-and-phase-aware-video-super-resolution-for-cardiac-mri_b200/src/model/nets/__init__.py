# Net registry: names resolved by `getattr(src.model.nets, config.net.name)` (reference src/main.py:59,128,179).
from .base_net import BaseNet
from .refine_net import RefineNet
from .edsr_net import EDSRNet
from .bicubic import Bicubic
from .drf_net import DRFNet, DRFSISRNet
