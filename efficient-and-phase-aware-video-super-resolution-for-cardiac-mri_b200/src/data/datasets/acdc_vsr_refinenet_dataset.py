"""Cine-MRI sequences for RefineNet (reference src/data/datasets/acdc_vsr_refinenet_dataset.py:10-89).

Output contract per item (what RefineNet and the runners consume):
    lr_imgs : list of T + 2U tensors (1, h, w)   - the cardiac cycle circularly padded by U warm-up frames a side
    hr_imgs : list of T tensors (1, s*h, s*w)
    pos_code: (T + 2U, 1) float32                - cardiac-phase code, NOT normalised
    index   : int
Training items are (sequence, target frame t) pairs: the T = num_frames frames ending at t; validation / test items
are whole cycles.  Unlike the reference, which re-opens both NIfTI volumes and the position-code pickle for every
item (:54-55,66-67), volumes and codes are decoded once and kept.
"""
import pickle
from pathlib import Path

import numpy as np

from ..transforms import compose
from .base_dataset import BaseDataset


def _load_nifti(path):
    """(H, W, C, T) array of a preprocessed 2d+1d volume; nibabel when installed, else the built-in NIfTI-1 reader."""
    try:
        import nibabel as nib
    except ImportError:
        from pvsr import nifti
        return nifti.read(path)
    return np.asarray(nib.load(str(path)).dataobj)


def window_slices(T, t, num_frames, num_updated_frames, train):
    """(lr_start, lr_end, hr_start, hr_end) into the cycle tiled three times (:74-87)."""
    U = num_updated_frames
    if train:
        end = t + T + 1
        start = end - num_frames
        return start - U, end + U, start, end
    return T - U, 2 * T + U, 0, T


def device_transform_plan(transforms, augments):
    from ..transforms import Normalize, ToTensor
    mean, std = 0.0, 1.0
    for step in transforms.steps:
        if isinstance(step, Normalize):
            if step.means is None or step.means.size != 1:
                raise TypeError('per-image statistics / multi-channel Normalize cannot be served from device memory')
            # float64 masters: the kernel rounds them to the volume's dtype exactly like Normalize does on the host
            mean, std = float(step.means.reshape(-1)[0]), float((step.stds + 1e-10).reshape(-1)[0])
        elif not isinstance(step, ToTensor):
            raise TypeError(f'transform {type(step).__name__} cannot be served from device memory')
    for step in augments.steps:
        if not hasattr(step, 'decide'):
            raise TypeError(f'augment {type(step).__name__} cannot be served from device memory')
    return float(mean), float(std), list(augments.steps)


class AcdcVSRRefineNetDataset(BaseDataset):
    def __init__(self, downscale_factor, transforms, pos_code_path, augments=None, num_frames=5,
                 num_updated_frames=0, **kwargs):
        super().__init__(**kwargs)
        if downscale_factor not in (2, 3, 4):
            raise ValueError(f'The downscale factor should be 2, 3, 4. Got {downscale_factor}.')
        self.downscale_factor = downscale_factor
        self.transforms = compose(transforms)
        self.augments = compose(augments)
        self.num_frames = num_frames
        self.num_updated_frames = num_updated_frames
        self.pos_code_path = pos_code_path
        self._volumes, self._pos_codes = {}, None

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

    def _pos_code(self, lr_path):
        if self._pos_codes is None:
            with open(self.pos_code_path, 'rb') as f:
                self._pos_codes = pickle.load(f)
        patient = Path(lr_path).parts[-1].split('.')[0].split('_')[0]
        return self._pos_codes[patient]

    def __len__(self):
        return len(self.data)

    # ---- device-resident serving (pvsr.device_loader.DeviceDataloader) -------------------------------------------
    def _sequences(self):
        if not hasattr(self, '_seq_paths'):
            self._seq_paths, self._seq_of = [], {}
            for entry in self.data:
                if entry[0] not in self._seq_of:
                    self._seq_of[entry[0]] = len(self._seq_paths)
                    self._seq_paths.append((entry[0], entry[1]))
        return self._seq_paths

    def sequence_table(self):
        """[(lr volume (H,W,1,T), hr volume, positional code (T,))] of every distinct cine sequence."""
        return [(self._volume(lr), self._volume(hr), np.asarray(self._pos_code(lr), dtype=np.float32))
                for lr, hr in self._sequences()]

    def transform_plan(self):
        """(mean, std, augment steps) of the configured chains, or TypeError when they cannot run as one gather."""
        return device_transform_plan(self.transforms, self.augments)

    def window(self, index):
        """(sequence number, lr_start, lr_end, hr_start, hr_end): the frame window of item `index` in cycle indices
        (taken modulo T by the consumer) - the same slices __getitem__ cuts out of the tiled lists."""
        entry = self.data[index]
        self._sequences()
        T = self._volume(entry[0]).shape[-1]
        train = self.type == 'train'
        return (self._seq_of[entry[0]],) + window_slices(T, entry[2] if train else 0, self.num_frames,
                                                         self.num_updated_frames, train)

    def __getitem__(self, index):
        entry = self.data[index]
        lr_vol, hr_vol = self._volume(entry[0]), self._volume(entry[1])
        T = lr_vol.shape[-1]
        frames = [lr_vol[..., t] for t in range(T)] + [hr_vol[..., t] for t in range(T)]
        if self.type == 'train':
            frames = self.augments(*frames)
        frames = [f.permute(2, 0, 1).contiguous() for f in self.transforms(*frames)]
        lr_imgs, hr_imgs = frames[:T] * 3, frames[T:] * 3
        pos_code = self.transforms(self._pos_code(entry[0]), normalize_tags=[False]).repeat(3).unsqueeze(1)
        a, b, c, d = window_slices(T, entry[2] if self.type == 'train' else 0, self.num_frames,
                                   self.num_updated_frames, self.type == 'train')
        return {'lr_imgs': lr_imgs[a:b], 'hr_imgs': hr_imgs[c:d], 'pos_code': pos_code[a:b], 'index': index}


class Dsb15VSRRefineNetDataset(AcdcVSRRefineNetDataset):
    """Named by configs/test/refine_net/exp*_dsb15.yaml:6 but absent from the reference's registry
    (src/data/datasets/__init__.py:1-8); nothing in the ACDC class is ACDC-specific, so it is the same reader pointed
    at the DSB15 directory."""
