"""Input pipeline: the built-in NIfTI reader, the host-side descriptor logic of DeviceDataloader (CPU)
and bit-exact equality of its batches with the host `Dataloader` path on the same seeded decisions (GPU)."""
import gzip
import json
import os
import pickle
import random

import numpy as np
import pytest
import torch


def _make_acdc_tree(root, kind, n_seq=3, T=9, lr=(12, 10), scale=4, dtype=np.int16, seed=0):
    """<root>/<kind>/{LR/X4,HR}/patientNNN/patientNNN_2d+1d_sequenceMM.nii.gz + the position-code pickle."""
    from pvsr import nifti
    rng = np.random.RandomState(seed)
    codes = {}
    for n in range(n_seq):
        patient = f'patient{n + 1:03d}'
        name = f'{patient}_2d+1d_sequence{n + 1:02d}.nii.gz'
        h, w = lr[0] + 2 * n, lr[1] + n           # ragged sizes between sequences
        for sub, (hh, ww) in ((f'LR/X{scale}', (h, w)), ('HR', (h * scale, w * scale))):
            d = root / kind / sub / patient
            d.mkdir(parents=True, exist_ok=True)
            vol = rng.randint(0, 255, size=(hh, ww, 1, T + n))
            nifti.write(d / name, vol.astype(dtype))
        codes[patient] = np.cos(np.linspace(0, np.pi, T + n)).astype(np.float64)
    with open(root / 'pos.pkl', 'wb') as f:
        pickle.dump(codes, f)
    return root / 'pos.pkl'


def _dataset(root, kind, pos, augment=True, **kw):
    from src.data.datasets import AcdcVSRRefineNetDataset
    augs = [dict(name='RandomHorizontalFlip'), dict(name='RandomVerticalFlip'),
            dict(name='RandomCropPatch', kwargs=dict(size=[8, 6], ratio=4))] if augment else None
    return AcdcVSRRefineNetDataset(data_dir=root, type=kind, downscale_factor=4, pos_code_path=pos,
                                   transforms=[dict(name='Normalize', kwargs=dict(means=[54.089], stds=[48.084])),
                                               dict(name='ToTensor')],
                                   augments=augs, num_frames=5, num_updated_frames=3, **kw)


def test_nifti_round_trip_and_layout(tmp_path):
    from pvsr import nifti
    for dtype in (np.uint8, np.int16, np.uint16, np.int32, np.float32, np.float64):
        a = (np.random.RandomState(1).rand(5, 7, 1, 4) * 200).astype(dtype)
        for name in ('v.nii', 'v.nii.gz'):
            nifti.write(tmp_path / name, a)
            b = nifti.read(tmp_path / name)
            assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)
    # on disk the FIRST index is the fastest one, after a 352-byte header
    a = np.arange(24, dtype=np.int16).reshape(2, 3, 4)
    nifti.write(tmp_path / 'o.nii', a)
    raw = open(tmp_path / 'o.nii', 'rb').read()
    assert len(raw) == 352 + 48 and raw[344:348] == b'n+1\0'
    assert np.array_equal(np.frombuffer(raw[352:], dtype='<i2'), a.ravel(order='F'))
    # scl_slope / scl_inter are applied like nibabel's dataobj does
    hdr = bytearray(raw)
    import struct
    struct.pack_into('<2f', hdr, 112, 2.0, 1.0)
    open(tmp_path / 's.nii', 'wb').write(bytes(hdr))
    assert np.array_equal(nifti.read(tmp_path / 's.nii'), a * 2.0 + 1.0)
    # big-endian files, bad magic, truncation
    be = bytearray(352)
    struct.pack_into('>i', be, 0, 348)
    struct.pack_into('>8h', be, 40, 2, 2, 2, 1, 1, 1, 1, 1)
    struct.pack_into('>2h', be, 70, 4, 16)
    struct.pack_into('>f', be, 108, 352.0)
    be[344:348] = b'n+1\0'
    open(tmp_path / 'b.nii', 'wb').write(bytes(be) + np.array([1, 2, 3, 4], dtype='>i2').tobytes())
    assert np.array_equal(nifti.read(tmp_path / 'b.nii'), np.array([[1, 3], [2, 4]]))
    with pytest.raises(nifti.NiftiError):
        open(tmp_path / 't.nii', 'wb').write(raw[:380])
        nifti.read(tmp_path / 't.nii')
    with pytest.raises(nifti.NiftiError):
        bad = bytearray(raw)
        bad[344:348] = b'xxxx'
        open(tmp_path / 'm.nii', 'wb').write(bytes(bad))
        nifti.read(tmp_path / 'm.nii')


def test_dataset_reads_nifti_tree_without_nibabel(tmp_path):
    pos = _make_acdc_tree(tmp_path, 'valid')
    ds = _dataset(tmp_path, 'valid', pos, augment=False)
    assert len(ds) == 3
    item = ds[1]
    T, U = 10, 3
    assert len(item['lr_imgs']) == T + 2 * U and len(item['hr_imgs']) == T
    assert item['lr_imgs'][0].shape == (1, 14, 11) and item['hr_imgs'][0].shape == (1, 56, 44)
    assert item['pos_code'].shape == (T + 2 * U, 1) and item['pos_code'].dtype == torch.float32
    # circular padding: frame U is phase 0, frame 0 is phase T-U
    assert torch.equal(item['lr_imgs'][0], item['lr_imgs'][T]) and torch.equal(item['lr_imgs'][U], item['lr_imgs'][U + T])
    raw = ds._volume(ds.data[1][0])
    want = (raw[:, :, 0, 0].astype(np.float32) - np.float32(54.089)) / (np.float32(48.084) + np.float32(1e-10))
    assert np.array_equal(item['lr_imgs'][U][0].numpy(), want)
    seq, a, b, c, d = ds.window(1)
    assert (seq, a, b, c, d) == (1, T - U, 2 * T + U, 0, T)


def test_host_loader_with_worker_processes(tmp_path):
    """num_workers > 0: every worker re-seeds numpy from the parent's state (reference dataloader.py:48-53)."""
    from src.data.dataloader import Dataloader
    pos = _make_acdc_tree(tmp_path, 'train', n_seq=2)
    dl = Dataloader(_dataset(tmp_path, 'train', pos), batch_size=4, shuffle=True, num_workers=2)
    batches = list(dl)
    assert len(batches) == len(dl) and batches[0]['lr_imgs'][0].shape == (4, 1, 8, 6)


def _emulate_gather(vol_thw, first, n, aff, mean, std):
    """numpy statement of pvsr_cine_gather for one sample (include/pvsr.h)."""
    T = vol_thw.shape[0]
    ys = aff.ay * np.arange(aff.h) + aff.by
    xs = aff.ax * np.arange(aff.w) + aff.bx
    if vol_thw.dtype == np.float64:    # fp64 volumes: fp64 arithmetic, one rounding (numpy semantics of the reference)
        out = [vol_thw[(first + f) % T][np.ix_(ys, xs)] for f in range(n)]
        return ((np.stack(out) - np.float64(mean)) / np.float64(std)).astype(np.float32)
    out = [vol_thw[(first + f) % T][np.ix_(ys, xs)].astype(np.float32) for f in range(n)]
    return (np.stack(out) - np.float32(mean)) / np.float32(std)


def test_descriptor_logic_matches_the_host_transform_chain(tmp_path):
    """The decisions DeviceDataloader draws, turned into affine gathers, reproduce the host items bit for bit."""
    from pvsr.device_loader import Affine
    pos = _make_acdc_tree(tmp_path, 'train')
    ds = _dataset(tmp_path, 'train', pos)
    mean, std, augs = ds.transform_plan()
    assert abs(mean - 54.089) < 1e-5 and abs(std - 48.084) < 1e-5 and len(augs) == 3
    table = ds.sequence_table()
    for index in (0, 7, 13, len(ds) - 1):
        random.seed(100 + index)
        item = ds[index]
        random.seed(100 + index)
        seq, a, b, c, d = ds.window(index)
        lrv, hrv, code = table[seq]
        lr_aff, hr_aff = Affine(*lrv.shape[:2]), Affine(*hrv.shape[:2])
        for aug in augs:
            step = aug.decide(lr_aff.h, lr_aff.w)
            if step is None:
                continue
            if step[0] == 'flip':
                lr_aff.flip(step[1]); hr_aff.flip(step[1])
            else:
                _, y0, x0, ph, pw, r = step
                lr_aff.crop(y0, x0, ph, pw); hr_aff.crop(y0 * r, x0 * r, ph * r, pw * r)
        lr = _emulate_gather(np.transpose(lrv[:, :, 0], (2, 0, 1)), a, b - a, lr_aff, mean, std)
        hr = _emulate_gather(np.transpose(hrv[:, :, 0], (2, 0, 1)), c, d - c, hr_aff, mean, std)
        assert np.array_equal(lr, torch.stack(item['lr_imgs'])[:, 0].numpy())
        assert np.array_equal(hr, torch.stack(item['hr_imgs'])[:, 0].numpy())
        T = code.shape[0]
        assert np.array_equal(np.array([code[(a + f) % T] for f in range(b - a)], dtype=np.float32),
                              item['pos_code'][:, 0].numpy())


# ------------------------------------------------------------------------------------------------ reference pin
GOLDEN_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_pipeline.npz")


def _load_data_golden():
    z = np.load(GOLDEN_DATA, allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def _golden_tree(root, z, meta):
    """The volumes the reference pipeline was run on (oracle/make_golden_data.py), written as NIfTI files."""
    from pvsr import nifti
    for key in z.files:
        if key.startswith("vol::"):
            path = root / key[len("vol::"):]
            path.parent.mkdir(parents=True, exist_ok=True)
            nifti.write(path, z[key])
    codes = {k[len("code::"):]: z[k] for k in z.files if k.startswith("code::")}
    with open(root / "pos.pkl", "wb") as f:
        pickle.dump(codes, f)
    return root / "pos.pkl"


def _golden_dataset(root, kind, pos, meta):
    from src.data.datasets import AcdcVSRRefineNetDataset
    return AcdcVSRRefineNetDataset(
        data_dir=root, type=kind, downscale_factor=meta["scale"], pos_code_path=pos,
        transforms=[dict(name='Normalize', kwargs=dict(means=meta["means"], stds=meta["stds"])), dict(name='ToTensor')],
        augments=[dict(name='RandomHorizontalFlip'), dict(name='RandomVerticalFlip'),
                  dict(name='RandomCropPatch', kwargs=dict(size=meta["patch"], ratio=meta["scale"]))],
        num_frames=meta["num_frames"], num_updated_frames=meta["num_updated_frames"])


def test_transforms_reproduce_the_reference_classes():
    """Normalize / ToTensor / RandomHorizontalFlip / RandomVerticalFlip / RandomCropPatch against outputs of the
    UNMODIFIED reference classes (src/data/transforms.py:74-168, 321-450) under the same `random.seed`: bit-exact."""
    from src.data import transforms as tr
    z, meta = _load_data_golden()
    lr, hr = list(z["t_lr"]), list(z["t_hr"])
    norm = tr.compose([dict(name='Normalize', kwargs=dict(means=meta["means"], stds=meta["stds"])), dict(name='ToTensor')])
    assert np.array_equal(torch.stack(norm(*lr)).numpy(), z["t_norm"])
    assert np.array_equal(torch.stack(norm(*[x.astype(np.float64) for x in lr])).numpy(), z["t_norm_f64"])
    per_image = tr.compose([dict(name='Normalize'), dict(name='ToTensor')])
    assert np.array_equal(torch.stack(per_image(*lr)).numpy(), z["t_norm_image_level"])
    code = norm(z["t_code_in"], normalize_tags=[False])
    assert code.dtype == torch.float32 and np.array_equal(code.numpy(), z["t_code_out"])
    chain = tr.compose([dict(name='RandomHorizontalFlip'), dict(name='RandomVerticalFlip'),
                        dict(name='RandomCropPatch', kwargs=dict(size=meta["patch"], ratio=meta["scale"]))])
    for seed in meta["aug_seeds"]:
        random.seed(seed)
        out = chain(*(lr + hr))
        assert np.array_equal(np.stack(out[:3]), z[f"t_aug_lr_{seed}"]), seed
        assert np.array_equal(np.stack(out[3:]), z[f"t_aug_hr_{seed}"]), seed


def test_host_dataset_reproduces_the_reference_items(tmp_path):
    """AcdcVSRRefineNetDataset.__getitem__ (train and whole-cycle items) against the dicts the UNMODIFIED reference
    dataset returned on the same volumes and `random.seed` (acdc_vsr_refinenet_dataset.py:49-89): bit-exact."""
    z, meta = _load_data_golden()
    pos = _golden_tree(tmp_path, z, meta)
    sets = {k: _golden_dataset(tmp_path, k, pos, meta) for k in ("train", "valid")}
    assert len(sets["train"]) == meta["len_train"] and len(sets["valid"]) == meta["len_valid"]
    for it in meta["items"]:
        random.seed(it["seed"])
        item = sets[it["kind"]][it["index"]]
        tag = f"{it['kind']}_{it['index']}"
        assert len(item["lr_imgs"]) == it["n_lr"] and len(item["hr_imgs"]) == it["n_hr"]
        assert list(item["lr_imgs"][0].shape) == it["lr_shape"] and list(item["hr_imgs"][0].shape) == it["hr_shape"]
        assert np.array_equal(torch.stack(item["lr_imgs"]).numpy(), z[f"item_lr::{tag}"]), tag
        assert np.array_equal(torch.stack(item["hr_imgs"]).numpy(), z[f"item_hr::{tag}"]), tag
        assert item["pos_code"].dtype == torch.float32
        assert np.array_equal(item["pos_code"].numpy(), z[f"item_pos::{tag}"]), tag
        assert item["index"] == it["index"]


def test_plain_vsr_dataset_reproduces_the_reference_items(tmp_path):
    """AcdcVSRDataset.__getitem__ ('last' and 'middle' windows incl. the wrap-around at both ends of the cycle, and
    whole-cycle items) against the dicts the UNMODIFIED reference dataset returned on the same volumes and `random.seed`
    (acdc_vsr_dataset.py:49-88): bit-exact."""
    from src.data.datasets import AcdcVSRDataset
    z, meta = _load_data_golden()
    _golden_tree(tmp_path, z, meta)
    v = np.load(os.path.join(os.path.dirname(GOLDEN_DATA), "data_vsr.npz"), allow_pickle=False)
    vmeta = json.loads(str(v["meta"]))
    sets = {}
    for it in vmeta["items"]:
        key = (it["order"], it["kind"])
        if key not in sets:
            sets[key] = AcdcVSRDataset(
                data_dir=tmp_path, type=it["kind"], downscale_factor=meta["scale"], temporal_order=it["order"],
                transforms=[dict(name='Normalize', kwargs=dict(means=meta["means"], stds=meta["stds"])), dict(name='ToTensor')],
                augments=[dict(name='RandomHorizontalFlip'), dict(name='RandomVerticalFlip'),
                          dict(name='RandomCropPatch', kwargs=dict(size=meta["patch"], ratio=meta["scale"]))],
                num_frames=vmeta["num_frames"])
            assert len(sets[key]) == vmeta[f"len_{it['kind']}"]
        random.seed(it["seed"])
        item = sets[key][it["index"]]
        tag = f"{it['order']}_{it['kind']}_{it['index']}"
        assert sorted(item) == ["hr_imgs", "index", "lr_imgs"] and item["index"] == it["index"]
        assert len(item["lr_imgs"]) == it["n"] == len(item["hr_imgs"])
        assert np.array_equal(torch.stack(item["lr_imgs"]).numpy(), v[f"item_lr::{tag}"]), tag
        assert np.array_equal(torch.stack(item["hr_imgs"]).numpy(), v[f"item_hr::{tag}"]), tag


@pytest.mark.gpu
def test_device_loader_reproduces_the_reference_items(tmp_path, pvsr_lib):
    """DeviceDataloader (volumes resident in HBM, one pvsr_cine_gather launch per resolution) against the same
    reference-recorded items: bit-exact, so the device path is pinned to the reference and not only to the host path."""
    from src.data.dataloader import DeviceDataloader
    z, meta = _load_data_golden()
    pos = _golden_tree(tmp_path, z, meta)
    loaders = {k: DeviceDataloader(_golden_dataset(tmp_path, k, pos, meta), batch_size=1) for k in ("train", "valid")}
    for it in meta["items"]:
        random.seed(it["seed"])
        batch = loaders[it["kind"]].fetch([it["index"]])
        tag = f"{it['kind']}_{it['index']}"
        lr = torch.stack([x[0] for x in batch["lr_imgs"]]).cpu().numpy()
        hr = torch.stack([x[0] for x in batch["hr_imgs"]]).cpu().numpy()
        assert np.array_equal(lr, z[f"item_lr::{tag}"]), tag
        assert np.array_equal(hr, z[f"item_hr::{tag}"]), tag
        assert np.array_equal(batch["pos_code"][0].cpu().numpy(), z[f"item_pos::{tag}"]), tag


def test_device_loader_refuses_cpu_and_unservable_chains(tmp_path):
    from pvsr.device_loader import DeviceDataloader
    from pvsr.lib import PvsrError
    pos = _make_acdc_tree(tmp_path, 'train', n_seq=1)
    if not torch.cuda.is_available():
        with pytest.raises(PvsrError):
            DeviceDataloader(_dataset(tmp_path, 'train', pos))
    from src.data.datasets import AcdcVSRRefineNetDataset
    ds = AcdcVSRRefineNetDataset(data_dir=tmp_path, type='train', downscale_factor=4, pos_code_path=pos,
                                 transforms=[dict(name='Normalize'), dict(name='ToTensor')], num_frames=5,
                                 num_updated_frames=3)
    with pytest.raises(TypeError):
        ds.transform_plan()          # per-image statistics cannot be a fixed (mean, std) gather


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.int16, np.float32, np.uint8, np.uint16, np.float64])
def test_device_batches_equal_host_batches(tmp_path, dtype, pvsr_lib):
    from src.data.dataloader import Dataloader, DeviceDataloader
    pos = _make_acdc_tree(tmp_path, 'train', dtype=dtype)
    pos = _make_acdc_tree(tmp_path, 'valid', dtype=dtype, seed=5)
    for kind, bs in (('train', 4), ('valid', 1)):
        host = Dataloader(_dataset(tmp_path, kind, pos), batch_size=bs, shuffle=False, num_workers=0)
        dev = DeviceDataloader(_dataset(tmp_path, kind, pos), batch_size=bs, shuffle=False, num_workers=8)
        assert len(host) == len(dev)
        random.seed(3)
        want = list(host)
        random.seed(3)
        got = list(dev)
        for w, g in zip(want, got):
            assert len(w['lr_imgs']) == len(g['lr_imgs']) and len(w['hr_imgs']) == len(g['hr_imgs'])
            for a, b in zip(w['lr_imgs'] + w['hr_imgs'], g['lr_imgs'] + g['hr_imgs']):
                assert b.is_cuda and a.shape == b.shape and torch.equal(a, b.cpu())
            assert torch.equal(w['pos_code'], g['pos_code'].cpu())
            assert torch.equal(w['index'], g['index'].cpu())
    assert dev.resident_bytes() > 0


@pytest.mark.gpu
def test_device_loader_feeds_trainer_and_predictor(tmp_path, pvsr_lib):
    """Same log from the host loader and the device loader for one validation epoch + the test loop."""
    from src.data.dataloader import Dataloader, DeviceDataloader
    from src.model.nets import RefineNet
    from src.model.metrics import PSNR
    from src.runner.predictors import AcdcVSRRefineNetPredictor
    pos = _make_acdc_tree(tmp_path, 'test', n_seq=3, lr=(12, 10))
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=2, update_memory=True,
              num_updated_frames=3, refine_window_size=5, upscale_factor=4, positional_encoding=True)
    logs = []
    for cls in (Dataloader, DeviceDataloader):
        torch.manual_seed(0)
        net = RefineNet(**kw)
        loader = cls(_dataset(tmp_path, 'test', pos, augment=False), batch_size=1, shuffle=False)
        pred = AcdcVSRRefineNetPredictor(device=torch.device('cuda:0'), test_dataloader=loader, net=net,
                                         loss_fns=[torch.nn.L1Loss()], loss_weights=[1.0], metric_fns=[PSNR()])
        logs.append(pred.predict())
    assert logs[0].keys() == logs[1].keys()
    for k in logs[0]:
        assert logs[0][k] == pytest.approx(logs[1][k], rel=1e-6), k


@pytest.mark.gpu
def test_device_loader_shards_cover_the_dataset_once(tmp_path, pvsr_lib):
    """shard=(rank, world): the ranks' batches partition the training items (DistributedSampler semantics)."""
    from src.data.dataloader import DeviceDataloader
    pos = _make_acdc_tree(tmp_path, 'train', n_seq=2)
    seen = []
    for rank in range(2):
        dl = DeviceDataloader(_dataset(tmp_path, 'train', pos), batch_size=3, shuffle=False, shard=(rank, 2))
        for batch in dl:
            assert batch['lr_imgs'][0].is_cuda
            seen.extend(batch['index'].tolist())
    n = len(_dataset(tmp_path, 'train', pos))
    assert sorted(set(seen)) == list(range(n)) and len(seen) in (n, n + 1)      # odd sizes are padded by one item
