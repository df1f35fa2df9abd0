"""Cine-MRI sequences for the plain video-SR nets - DRFNet on this path (reference
src/data/datasets/acdc_vsr_dataset.py:8-88, dsb15_vsr_dataset.py).

Output contract per item:
    lr_imgs : list of n tensors (1, h, w)          hr_imgs : list of n tensors (1, s*h, s*w)          index : int
Training items are (sequence, target frame t) pairs: n = num_frames frames, `temporal_order` 'last' = {t-n+1 .. t},
'middle' = {t-(n-1)//2 .. t+(n-1)-(n-1)//2}, wrapped around the cardiac cycle (:58-73); validation / test items are
whole cycles.  Unlike the reference, which re-opens both NIfTI volumes for every item (:54-55), volumes are decoded
once and kept.
"""
from pathlib import Path

from ..transforms import compose
from .acdc_vsr_refinenet_dataset import _load_nifti
from .base_dataset import BaseDataset


def frame_window(T, t, num_frames, temporal_order):
    """Cycle indices of a training item (:58-73).  The reference concatenates `[..., start:]` and `[..., :end]` when the
    window leaves [0, T) - for num_frames <= T that is (start + i) mod T."""
    n = num_frames
    if n > T:
        raise ValueError(f'num_frames ({n}) exceeds the length of the sequence ({T}).')
    start = t - n + 1 if temporal_order == 'last' else t - (n - 1) // 2
    return [(start + i) % T for i in range(n)]


class AcdcVSRDataset(BaseDataset):
    def __init__(self, downscale_factor, transforms, augments=None, num_frames=5, temporal_order='last', **kwargs):
        super().__init__(**kwargs)
        if downscale_factor not in (2, 3, 4):
            raise ValueError(f'The downscale factor should be 2, 3, 4. Got {downscale_factor}.')
        if temporal_order not in ('last', 'middle'):
            raise ValueError(f"The temporal order should be 'last' or 'middle'. Got {temporal_order}.")
        self.downscale_factor = downscale_factor
        self.transforms = compose(transforms)
        self.augments = compose(augments)
        self.num_frames = num_frames
        self.temporal_order = temporal_order
        self._volumes = {}

        root = Path(self.data_dir) / self.type
        lr_paths = sorted((root / 'LR' / f'X{downscale_factor}').glob('**/*2d+1d*.nii.gz'))
        hr_paths = sorted((root / 'HR').glob('**/*2d+1d*.nii.gz'))
        if self.type == 'train':
            self.data = []
            for lr_path, hr_path in zip(lr_paths, hr_paths):
                T = self._volume(lr_path).shape[-1]
                self.data.extend((lr_path, hr_path, t) for t in range(T))
        else:
            self.data = list(zip(lr_paths, hr_paths))

    def _volume(self, path):
        if path not in self._volumes:
            self._volumes[path] = _load_nifti(path)      # (H, W, C, T)
        return self._volumes[path]

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        entry = self.data[index]
        lr_vol, hr_vol = self._volume(entry[0]), self._volume(entry[1])
        T = lr_vol.shape[-1]
        train = self.type == 'train'
        frames = frame_window(T, entry[2], self.num_frames, self.temporal_order) if train else list(range(T))
        imgs = [lr_vol[..., t] for t in frames] + [hr_vol[..., t] for t in frames]        # list of (H, W, C)
        if train:
            imgs = self.augments(*imgs)
        imgs = [img.permute(2, 0, 1).contiguous() for img in self.transforms(*imgs)]
        return {'lr_imgs': imgs[:len(imgs) // 2], 'hr_imgs': imgs[len(imgs) // 2:], 'index': index}


class Dsb15VSRDataset(AcdcVSRDataset):
    """The same reader pointed at the DSB15 directory (reference dsb15_vsr_dataset.py differs in its docstring only)."""
