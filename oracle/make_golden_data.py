"""ORACLE TOOLING - pins the INPUT PIPELINE to the reference: runs the UNMODIFIED reference
`src/data/transforms.py` (Normalize :100-168, ToTensor :74-97, RandomHorizontalFlip / RandomVerticalFlip :321-376,
RandomCropPatch :379-450) and `src/data/datasets/acdc_vsr_refinenet_dataset.py` (`__getitem__` :49-89) in the build
container and records what they return -> tests/golden/data_pipeline.npz; the items of the plain video-SR dataset
(`src/data/datasets/acdc_vsr_dataset.py` :49-88, both temporal orders) on the same volumes -> tests/golden/data_vsr.npz.

    python oracle/make_golden_data.py

The two modules import SimpleITK and nibabel (not installed).  Only `RandomElasticDeformation` touches SimpleITK and
only `nib.load` touches nibabel, so both are stubbed in sys.modules: `nib.load(path)` returns an object whose
`.get_data()` is the seeded float32 volume registered for that file name (acdc_preprocess.py:40 saves float32) and whose
`.header.get_data_shape()` is its shape.  `Box` below stands in for python-box (attribute access + .get on the YAML
entries, transforms.py:21-24).  Nothing else deviates from the reference; its random decisions come from Python's
`random` module (transforms.py:340,369,443), seeded here per record with `random.seed(seed)`.

The tests rebuild the same volumes from the stored arrays (as NIfTI files written by pvsr.nifti) and require the
product's host `Dataloader` path AND the HBM-resident `DeviceDataloader` to reproduce every recorded item bit-exactly.
"""
import importlib
import json
import os
import pickle
import random
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

class Box(dict):
    __getattr__ = dict.get


class _FakeImage:
    def __init__(self, arr):
        self._arr = arr
        self.header = types.SimpleNamespace(get_data_shape=lambda: arr.shape)

    def get_data(self):
        return self._arr


def load_reference_data_modules():
    sys.path.insert(0, REF)
    sitk = types.ModuleType("SimpleITK")
    nib = types.ModuleType("nibabel")
    nib.load = None     # bound per split in main(): serves VOLUMES_BY_KEY by relative path
    sys.modules["SimpleITK"], sys.modules["nibabel"] = sitk, nib
    for n, p in [("src", REF + "/src"), ("src.data", REF + "/src/data"), ("src.data.datasets", REF + "/src/data/datasets")]:
        m = types.ModuleType(n)
        m.__path__ = [p]
        sys.modules[n] = m
    sys.modules["src"].data = sys.modules["src.data"]          # what a real package import would have bound
    sys.modules["src.data"].datasets = sys.modules["src.data.datasets"]
    tr = importlib.import_module("src.data.transforms")
    ds = importlib.import_module("src.data.datasets.acdc_vsr_refinenet_dataset")
    return tr, ds


VOLUMES_BY_KEY = {}   # "<split>/<LR/X4|HR>/<patient>/<file>" -> (H, W, 1, T) float32 array served by the nibabel stub

SEQS = [  # (patient, sequence id, LR h, LR w, T)
    ("patient001", 1, 12, 10, 9), ("patient002", 2, 14, 11, 10), ("patient003", 3, 16, 12, 11)]
SCALE, NUM_FRAMES, U, PATCH = 4, 5, 3, (8, 6)


def make_tree(root):
    """Empty files with the names the dataset globs for + the volumes behind them + the position-code pickle."""
    rng = np.random.RandomState(2024)
    codes = {}
    for kind in ("train", "valid"):
        for patient, sid, h, w, T in SEQS:
            name = f"{patient}_2d+1d_sequence{sid:02d}.nii.gz"
            for sub, (hh, ww) in ((f"LR/X{SCALE}", (h, w)), ("HR", (h * SCALE, w * SCALE))):
                d = Path(root) / kind / sub / patient
                d.mkdir(parents=True, exist_ok=True)
                (d / name).touch()
                # integer-valued float32 in [0, 255] like the preprocessed volumes (acdc_preprocess.py:34-40)
                key = f"{kind}/{sub}/{patient}/{name}"
                vol = rng.randint(0, 256, size=(hh, ww, 1, T)).astype(np.float32)
                VOLUMES_BY_KEY[key] = vol
            codes[patient] = np.cos(np.linspace(0, 2 * np.pi, T, endpoint=False))     # float64, like gen_positional_encoding
    pos = Path(root) / "pos.pkl"
    with open(pos, "wb") as f:
        pickle.dump(codes, f)
    return pos, codes


def main():
    tr, ds = load_reference_data_modules()
    rec = {}
    meta = {"scale": SCALE, "num_frames": NUM_FRAMES, "num_updated_frames": U, "patch": list(PATCH), "seqs": SEQS,
            "means": [54.089], "stds": [48.084], "items": []}

    # ---- transform level: every step on its own, one seed per record
    g = np.random.RandomState(7)
    lr = [g.randint(0, 256, size=(12, 10, 1)).astype(np.float32) for _ in range(3)]
    hr = [g.randint(0, 256, size=(48, 40, 1)).astype(np.float32) for _ in range(3)]
    rec["t_lr"], rec["t_hr"] = np.stack(lr), np.stack(hr)
    norm = tr.Compose([tr.Normalize(means=[54.089], stds=[48.084]), tr.ToTensor()])
    rec["t_norm"] = torch.stack(norm(*lr)).numpy()
    f64 = [x.astype(np.float64) for x in lr]
    rec["t_norm_f64"] = torch.stack(norm(*f64)).numpy()          # computed in float64, then ToTensor's .float()
    rec["t_norm_image_level"] = torch.stack(tr.Compose([tr.Normalize(), tr.ToTensor()])(*lr)).numpy()
    code = np.cos(np.linspace(0, 2 * np.pi, 9, endpoint=False))
    rec["t_code_in"] = code
    rec["t_code_out"] = norm(code, normalize_tags=[False]).numpy()                 # untouched, float64 -> float32
    chain = tr.Compose([tr.RandomHorizontalFlip(), tr.RandomVerticalFlip(), tr.RandomCropPatch(size=list(PATCH), ratio=SCALE)])
    for seed in range(12):
        random.seed(seed)
        out = chain(*(lr + hr))
        rec[f"t_aug_lr_{seed}"] = np.stack([np.ascontiguousarray(o) for o in out[:3]])
        rec[f"t_aug_hr_{seed}"] = np.stack([np.ascontiguousarray(o) for o in out[3:]])
    meta["aug_seeds"] = list(range(12))

    # ---- dataset level
    with tempfile.TemporaryDirectory() as td:
        pos_path, codes = make_tree(td)
        for k, v in VOLUMES_BY_KEY.items():
            rec["vol::" + k] = v
        for p, c in codes.items():
            rec["code::" + p] = c
        cfg_t = [Box(name="Normalize", kwargs=Box(means=[54.089], stds=[48.084])), Box(name="ToTensor")]
        cfg_a = [Box(name="RandomHorizontalFlip"), Box(name="RandomVerticalFlip"),
                 Box(name="RandomCropPatch", kwargs=Box(size=list(PATCH), ratio=SCALE))]
        for kind in ("train", "valid"):
            # the LR and the HR file of a sequence share one file name: serve them by full relative path
            def fake_load(path, kind=kind):
                parts = Path(path).parts
                i = parts.index(kind)
                return _FakeImage(VOLUMES_BY_KEY["/".join(parts[i:])])
            sys.modules["nibabel"].load = fake_load
            dset = ds.AcdcVSRRefineNetDataset(downscale_factor=SCALE, transforms=cfg_t, pos_code_path=str(pos_path),
                                              augments=cfg_a, num_frames=NUM_FRAMES, num_updated_frames=U,
                                              data_dir=Path(td), type=kind)
            meta[f"len_{kind}"] = len(dset)
            picks = [0, 4, 8, 9, 17, len(dset) - 1] if kind == "train" else list(range(len(dset)))
            for j, index in enumerate(picks):
                seed = 1000 + 7 * j + (0 if kind == "train" else 500)
                random.seed(seed)
                item = dset[index]
                tag = f"{kind}_{index}"
                rec[f"item_lr::{tag}"] = torch.stack(item["lr_imgs"]).numpy()
                rec[f"item_hr::{tag}"] = torch.stack(item["hr_imgs"]).numpy()
                rec[f"item_pos::{tag}"] = item["pos_code"].numpy()
                assert item["index"] == index
                meta["items"].append({"kind": kind, "index": index, "seed": seed,
                                      "n_lr": len(item["lr_imgs"]), "n_hr": len(item["hr_imgs"]),
                                      "lr_shape": list(item["lr_imgs"][0].shape), "hr_shape": list(item["hr_imgs"][0].shape)})
        # ---- the plain video-SR dataset (acdc_vsr_dataset.py:49-88) on the SAME volumes -> data_vsr.npz (items only)
        vds = importlib.import_module("src.data.datasets.acdc_vsr_dataset")
        vrec, vmeta = {}, {"num_frames": NUM_FRAMES, "items": []}
        for order in ("last", "middle"):
            for kind in ("train", "valid"):
                def fake_load(path, kind=kind):
                    parts = Path(path).parts
                    i = parts.index(kind)
                    return _FakeImage(VOLUMES_BY_KEY["/".join(parts[i:])])
                sys.modules["nibabel"].load = fake_load
                dset = vds.AcdcVSRDataset(downscale_factor=SCALE, transforms=cfg_t, augments=cfg_a, num_frames=NUM_FRAMES,
                                          temporal_order=order, data_dir=Path(td), type=kind)
                vmeta[f"len_{kind}"] = len(dset)
                picks = [0, 1, 3, 7, 8, 9, 18, len(dset) - 1] if kind == "train" else list(range(len(dset)))
                for j, index in enumerate(picks):
                    seed = 3000 + 11 * j + (0 if kind == "train" else 500) + (0 if order == "last" else 97)
                    random.seed(seed)
                    item = dset[index]
                    tag = f"{order}_{kind}_{index}"
                    vrec[f"item_lr::{tag}"] = torch.stack(item["lr_imgs"]).numpy()
                    vrec[f"item_hr::{tag}"] = torch.stack(item["hr_imgs"]).numpy()
                    assert item["index"] == index and sorted(item) == ["hr_imgs", "index", "lr_imgs"]
                    vmeta["items"].append({"order": order, "kind": kind, "index": index, "seed": seed,
                                           "n": len(item["lr_imgs"])})
        vrec["meta"] = np.array(json.dumps(vmeta))
        vpath = os.path.join(OUT, "data_vsr.npz")
        np.savez_compressed(vpath, **vrec)
        print("wrote", vpath, os.path.getsize(vpath), "bytes;", len(vmeta["items"]), "plain-VSR dataset items")
    rec["meta"] = np.array(json.dumps(meta))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "data_pipeline.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path), "bytes;", len(meta["items"]), "dataset items")


if __name__ == "__main__":
    main()
