# Dataset registry: names resolved by `getattr(src.data.datasets, config.dataset.name)` (reference src/main.py:46).
from .base_dataset import BaseDataset
from .acdc_vsr_refinenet_dataset import AcdcVSRRefineNetDataset, Dsb15VSRRefineNetDataset
from .synthetic_cine_dataset import SyntheticCineDataset
from .acdc_sisr_dataset import AcdcSISRDataset, Dsb15SISRDataset, SyntheticSISRDataset
from .acdc_vsr_dataset import AcdcVSRDataset, Dsb15VSRDataset

__all__ = ['BaseDataset', 'AcdcVSRRefineNetDataset', 'Dsb15VSRRefineNetDataset', 'SyntheticCineDataset',
           'AcdcSISRDataset', 'Dsb15SISRDataset', 'SyntheticSISRDataset', 'AcdcVSRDataset', 'Dsb15VSRDataset']
