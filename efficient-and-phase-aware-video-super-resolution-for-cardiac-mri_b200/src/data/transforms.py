"""The preprocessing / augmentation steps named by the RefineNet configs
(configs/train/refine_net/exp1_x4.yaml:10-23): Normalize, ToTensor, RandomHorizontalFlip, RandomVerticalFlip,
RandomCropPatch, resolved by name through `compose` (reference src/data/transforms.py:10-28).

Every transform takes any number of numpy images (H, W, C) - the LR frames followed by the HR frames of one cine
sequence - and returns the same number, applying ONE random decision to all of them.

Parity with the reference is pinned by tests/golden/data_pipeline.npz (oracle/make_golden_data.py runs the unmodified
reference classes): the random decisions are drawn from Python's `random` module with the reference's calls in the
reference's order (`random.random() < prob` per flip, transforms.py:340,369; `random.randint(0, h - ht)` then
`random.randint(0, w - wt)` per crop, :443), and Normalize computes in the image's own floating dtype like numpy does
for `img[..., c] = (img[..., c] - mean) / (std + 1e-10)` (:165-168) before ToTensor's `.float()`.
"""
import importlib
import random

import numpy as np
import torch


def compose(transforms=None):
    """Builds the chain described by a list of {name, kwargs} entries; None -> identity chain."""
    if transforms is None:
        return Compose([])
    module = importlib.import_module(__name__)
    steps = []
    for t in transforms:
        cls = getattr(module, t['name'] if isinstance(t, dict) else t.name)
        kw = (t.get('kwargs') if isinstance(t, dict) else t.get('kwargs')) or {}
        steps.append(cls(**kw))
    return Compose(steps)


class BaseTransform:
    def __call__(self, *imgs, **kwargs):
        raise NotImplementedError


class Compose(BaseTransform):
    def __init__(self, steps):
        self.steps = steps

    def __call__(self, *imgs, **kwargs):
        for step in self.steps:
            imgs = step(*imgs, **kwargs)
            if not isinstance(imgs, tuple):
                imgs = (imgs,)
        return imgs[0] if len(imgs) == 1 else imgs


class ToTensor(BaseTransform):
    """numpy -> torch; `dtypes` optionally gives one torch dtype per image (default float32)."""

    def __call__(self, *imgs, dtypes=None, **kwargs):
        if dtypes is None:
            dtypes = [torch.float32] * len(imgs)
        return tuple(torch.as_tensor(np.ascontiguousarray(img)).to(dt) for img, dt in zip(imgs, dtypes))


class Normalize(BaseTransform):
    """(x - mean) / (std + 1e-10) per channel; without means/stds, per-image statistics.  `normalize_tags` selects
    which images are touched (the positional code is passed with [False], acdc_vsr_refinenet_dataset.py:71)."""

    def __init__(self, means=None, stds=None):
        if (means is None) != (stds is None):
            raise ValueError('means and stds should be both None or both given.')
        if means is not None and len(means) != len(stds):
            raise ValueError('The number of the means should be the same as the standard deviations.')
        # float64 masters (the YAML numbers); cast to the image's dtype at use, which is what numpy does with the
        # reference's Python-float operands
        self.means = None if means is None else np.asarray(means, dtype=np.float64)
        self.stds = None if stds is None else np.asarray(stds, dtype=np.float64)

    def __call__(self, *imgs, normalize_tags=None, **kwargs):
        tags = normalize_tags if normalize_tags is not None else [True] * len(imgs)
        out = []
        for img, tag in zip(imgs, tags):
            if tag:
                img = np.asarray(img)
                if not np.issubdtype(img.dtype, np.floating):
                    # the preprocessing scripts only write float32 volumes (acdc_preprocess.py:40); the reference would
                    # truncate the normalised values back into an integer array - integer volumes are promoted instead
                    img = img.astype(np.float32)
                dt = img.dtype
                if self.means is None:
                    axes = tuple(range(img.ndim - 1))
                    img = ((img - img.mean(axis=axes)) / (img.std(axis=axes) + 1e-10)).astype(dt, copy=False)
                else:
                    img = ((img - self.means.astype(dt)) / (self.stds + 1e-10).astype(dt)).astype(dt, copy=False)
            out.append(img)
        return tuple(out)


class _RandomFlip(BaseTransform):
    axis = 0

    def __init__(self, prob=0.5):
        self.prob = max(0, min(prob, 1))

    def decide(self, h, w):
        """Draws this step's decision for an (h, w) LR image: ('flip', axis) or None (pvsr.device_loader)."""
        return ('flip', self.axis) if random.random() < self.prob else None

    def __call__(self, *imgs, **kwargs):
        if self.decide(*imgs[0].shape[:2]) is not None:
            return tuple(np.flip(img, self.axis).copy() for img in imgs)
        return imgs


class RandomHorizontalFlip(_RandomFlip):
    axis = 1


class RandomVerticalFlip(_RandomFlip):
    axis = 0


class RandomCropPatch(BaseTransform):
    """One random LR window of `size` and the aligned `ratio`x larger HR window: the first half of the images are
    LR frames, the second half the HR frames of the same sequence."""

    def __init__(self, size, ratio):
        self.size, self.ratio = tuple(size), ratio

    def decide(self, h, w):
        """('crop', y0, x0, ph, pw, ratio) for an (h, w) LR image."""
        ph, pw = self.size
        if ph > h or pw > w:
            raise ValueError(f'The crop size {self.size} exceeds the LR image size {(h, w)}.')
        y0 = random.randint(0, h - ph)       # inclusive bounds, rows first (transforms.py:443)
        x0 = random.randint(0, w - pw)
        return ('crop', y0, x0, ph, pw, self.ratio)

    def __call__(self, *imgs, **kwargs):
        half = len(imgs) // 2
        _, y0, x0, ph, pw, r = self.decide(*imgs[0].shape[:2])
        lr = [img[y0:y0 + ph, x0:x0 + pw] for img in imgs[:half]]
        hr = [img[y0 * r:(y0 + ph) * r, x0 * r:(x0 + pw) * r] for img in imgs[half:]]
        return tuple(lr + hr)
