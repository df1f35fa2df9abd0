"""Single frames for the SISR nets (reference src/data/datasets/acdc_sisr_dataset.py:7-42; the DSB15 variant differs
only in its directory).  Item contract: {'lr_img': (1, h, w), 'hr_img': (1, s*h, s*w), 'index': int}; training items
are augmented as an (LR, HR) pair before the transforms.  Volumes are decoded once and kept (the reference re-opens
both NIfTI files per item, :36-37)."""
from pathlib import Path

import torch

from ..transforms import compose
from .acdc_vsr_refinenet_dataset import _load_nifti
from .base_dataset import BaseDataset


class AcdcSISRDataset(BaseDataset):
    def __init__(self, downscale_factor, transforms, augments=None, **kwargs):
        super().__init__(**kwargs)
        if downscale_factor not in (2, 3, 4):
            raise ValueError(f'The downscale factor should be 2, 3, 4. Got {downscale_factor}.')
        self.downscale_factor = downscale_factor
        self.transforms, self.augments = compose(transforms), compose(augments)
        root = Path(self.data_dir) / self.type
        lr_paths = sorted((root / 'LR' / f'X{downscale_factor}').glob('**/*2d*.nii.gz'))
        hr_paths = sorted((root / 'HR').glob('**/*2d*.nii.gz'))
        # the VSR volumes ('..._2d+1d_...') live in the same tree and match the reference's glob too (:27-28)
        self.data = list(zip(lr_paths, hr_paths))
        self._cache = {}

    def _image(self, path):
        if path not in self._cache:
            self._cache[path] = _load_nifti(path)          # (H, W, C)
        return self._cache[path]

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        lr_path, hr_path = self.data[index]
        lr_img, hr_img = self._image(lr_path), self._image(hr_path)
        if self.type == 'train':
            lr_img, hr_img = self.augments(lr_img, hr_img)
        lr_img = self.transforms(lr_img).permute(2, 0, 1).contiguous()
        hr_img = self.transforms(hr_img).permute(2, 0, 1).contiguous()
        return {'lr_img': lr_img, 'hr_img': hr_img, 'index': index}


class Dsb15SISRDataset(AcdcSISRDataset):
    """reference src/data/datasets/dsb15_sisr_dataset.py: the same reader pointed at the DSB15 directory."""


class SyntheticSISRDataset(BaseDataset):
    """ACDCSR-shaped random frames with the SISR item contract (the datasets are not available offline)."""

    def __init__(self, downscale_factor=4, num_images=64, lr_size=(54, 63), seed=1234, data_dir=None, type='test',
                 **_):
        super().__init__(data_dir=data_dir, type=type)
        self.downscale_factor, self.lr_size, self.seed = downscale_factor, tuple(lr_size), seed
        self.data = [(Path(f'synthetic{n // 30 + 1:03d}_2d_slice01_frame{n % 30 + 1:02d}.nii.gz'), None)
                     for n in range(num_images)]

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self.seed + index)
        (h, w), s = self.lr_size, self.downscale_factor
        return {'lr_img': torch.randn(1, h, w, generator=g), 'hr_img': torch.randn(1, h * s, w * s, generator=g),
                'index': index}
