#!/usr/bin/env python
"""Input pipeline: host `Dataloader` (the reference's numpy path: per-item transforms, default
collate, then H2D) against `DeviceDataloader` (volumes resident in HBM, one pvsr_cine_gather launch per resolution
per batch) on a synthetic ACDC-shaped NIfTI tree.  Both yield bit-identical batches (tests/test_device_loader.py).

  python profiles/bench_loader.py [--sequences 24] [--batches 40]
"""
import argparse
import json
import os
import pickle
import sys
import tempfile
import time
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
for p in (PKG, ROOT, os.path.join(PKG, "csrc")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def make_tree(root, kind, n_seq, T=30, lr=(54, 64), scale=4):
    from pvsr import nifti
    rng = np.random.RandomState(0)
    codes = {}
    for n in range(n_seq):
        patient = f'patient{n + 1:03d}'
        name = f'{patient}_2d+1d_sequence01.nii.gz'
        for sub, (hh, ww) in ((f'LR/X{scale}', lr), ('HR', (lr[0] * scale, lr[1] * scale))):
            d = root / kind / sub / patient
            d.mkdir(parents=True, exist_ok=True)
            nifti.write(d / name, rng.randint(0, 255, size=(hh, ww, 1, T)).astype(np.int16))
        codes[patient] = np.cos(np.linspace(0, np.pi, T))
    with open(root / 'pos.pkl', 'wb') as f:
        pickle.dump(codes, f)
    return root / 'pos.pkl'


def dataset(root, kind, pos):
    from src.data.datasets import AcdcVSRRefineNetDataset
    augs = [dict(name='RandomHorizontalFlip'), dict(name='RandomVerticalFlip'),
            dict(name='RandomCropPatch', kwargs=dict(size=[32, 32], ratio=4))]
    ds = AcdcVSRRefineNetDataset(data_dir=root, type=kind, downscale_factor=4, pos_code_path=pos,
                                 transforms=[dict(name='Normalize', kwargs=dict(means=[54.089], stds=[48.084])),
                                             dict(name='ToTensor')],
                                 augments=augs, num_frames=7, num_updated_frames=6)
    for e in ds.data:                     # decode every volume once up front for BOTH paths (not what is measured)
        ds._volume(e[0]); ds._volume(e[1])
    return ds


def run(loader, n_batches, dev):
    from src.runner.trainers.base_trainer import to_device
    it = iter(loader)
    b = to_device(next(it), dev)          # warm-up (worker start-up, residency upload)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = frames = 0
    for batch in it:
        batch = to_device(batch, dev)
        frames += len(batch['hr_imgs']) * batch['hr_imgs'][0].shape[0]
        n += 1
        if n == n_batches:
            break
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0), frames / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=24)
    ap.add_argument("--batches", type=int, default=40)
    args = ap.parse_args()
    import build as pvsr_build
    pvsr_build.build()
    from src.data.dataloader import Dataloader, DeviceDataloader
    dev = torch.device('cuda', 0)
    out = {"tree": f"{args.sequences} ACDC-shaped int16 sequences per split, T=30, LR 54x64, HR 216x256",
           "train_batch": "16 items: 19 LR patches 32x32 + 7 HR patches 128x128 each (flips + crop + normalise)",
           "test_batch": "1 whole cycle: 42 LR frames + 30 HR frames"}
    with tempfile.TemporaryDirectory() as tmp:
        root = Path(tmp)
        pos = make_tree(root, 'train', args.sequences)
        make_tree(root, 'test', args.sequences)
        for kind, bs in (('train', 16), ('test', 1)):
            res = {}
            for name, make in (("host_workers0", lambda d: Dataloader(d, batch_size=bs, shuffle=True, num_workers=0)),
                               ("host_workers8", lambda d: Dataloader(d, batch_size=bs, shuffle=True, num_workers=8,
                                                                      pin_memory=True)),
                               ("device", lambda d: DeviceDataloader(d, batch_size=bs, shuffle=True))):
                ds = dataset(root, kind, pos)
                nb = min(args.batches, len(ds) // bs - 1)
                bps, fps = run(make(ds), nb, dev)
                res[name] = {"batches_per_s": round(bps, 1), "target_frames_per_s": round(fps, 1)}
            out[kind] = res
        # the gather kernel alone: 32 whole cycles per launch pair, device-timed
        ds = dataset(root, 'test', pos)
        dl = DeviceDataloader(ds, batch_size=1)
        idx = list(range(min(32, len(ds))))
        for _ in range(3):
            dl.fetch(idx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            b = dl.fetch(idx)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        px = len(idx) * (42 * 54 * 64 + 30 * 216 * 256)
        out["gather_32_cycles"] = {"ms": ms, "GB_per_s": px * (2 + 4) / ms / 1e6,
                                   "bytes": "2 B read (int16) + 4 B written per pixel; includes the host-side "
                                            "descriptor build + 2 launches per fetch"}
        out["resident_MB"] = dl.resident_bytes() / 1e6
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
