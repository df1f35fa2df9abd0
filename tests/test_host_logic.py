"""CPU tests of the host side: the reference-facing `src` mirror (config registry, dataset contract, transforms,
metrics against values produced by the reference's own metrics.py, monitor, checkpoint layout), the sharding helpers,
and that libpvsr.so exports every symbol include/pvsr.h declares.  No compute call needs a GPU here."""
import json
import math
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol(pvsr_lib):
    hdr = open(os.path.join(ROOT, "include", "pvsr.h")).read()
    declared = set(re.findall(r"\b(pvsr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    from pvsr import lib as L
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    for name in declared:
        assert hasattr(pvsr_lib, name), name
    assert pvsr_lib.pvsr_version() == 100


def test_plan_creation_is_host_only_and_validates(pvsr_lib):
    """Plan geometry / accounting are pure host logic: FLOP totals equal SURVEY.md section 8d."""
    import ctypes as C
    from pvsr import lib as L

    def make(**kw):
        cfg = L.NetConfig()
        base = dict(batch=1, n_frames=42, n_updated=6, h=54, w=63, scale=4, n_stages=3, window=5, n_layers=3,
                    pos_enc=1, memory=1, all_heads=1, save_for_backward=0)
        base.update(kw)
        for k, v in base.items():
            setattr(cfg, k, v)
        h = C.c_void_p()
        return pvsr_lib.pvsr_plan_create(C.byref(cfg), C.byref(h)), h

    rc, h = make()
    assert rc == 0
    # reference-executed FLOPs of one ACDC x4 sequence: 3.527 TFLOP (all 9 heads), minus the last stage's dead work
    # the plan skips (results unchanged): (U - half) = 4 trailing ConvLSTM steps per direction and the 2 x 4 refine
    # windows outside the T output frames
    px = 54 * 63
    dead = 4 * 2 * 3 * 589824 * px + 8 * (1497690 + 148608) * px
    assert abs(dead / 1e12 - 0.093) < 1e-3
    assert abs(pvsr_lib.pvsr_plan_flops(h) / 1e12 - (3.527 - dead / 1e12)) < 0.01
    assert pvsr_lib.pvsr_plan_num_lists(h) == 9
    pvsr_lib.pvsr_plan_destroy(h)
    rc, h = make(all_heads=0)
    assert rc == 0 and abs(pvsr_lib.pvsr_plan_flops(h) / 1e12 - (2.308 - dead / 1e12)) < 0.01
    assert pvsr_lib.pvsr_plan_num_lists(h) == 1
    pvsr_lib.pvsr_plan_destroy(h)
    rc, h = make(batch=16, n_frames=19, h=32, w=32, save_for_backward=1)
    assert rc == 0
    dead_tr = (4 * 2 * 3 * 589824 + 8 * (1497690 + 148608)) * 32 * 32 * 16
    assert abs(pvsr_lib.pvsr_plan_flops(h) / 1e12 - (6.060 - dead_tr / 1e12)) < 0.01   # training forward
    assert abs(pvsr_lib.pvsr_plan_flops_bwd(h) / 1e12 - 6.65) < 0.06        # dgrad + wgrad on the gradient frames
    assert pvsr_lib.pvsr_plan_num_launches_bwd(h) > 0
    pvsr_lib.pvsr_plan_destroy(h)
    for bad in (dict(scale=5), dict(n_updated=0), dict(window=4), dict(n_layers=4), dict(n_frames=12),
                dict(save_for_backward=1, all_heads=0)):
        rc, _ = make(**bad)
        assert rc < 0 and pvsr_lib.pvsr_last_error()


def test_pack_index_is_a_permutation_of_the_weight(pvsr_lib):
    from pvsr import ops
    for spec, shape in ((ops.spec_lstm(), (256, 128, 3, 3)), (ops.spec_refine_conv2(), (64, 129, 3, 3)),
                        (ops.spec_head_ps(2), (256, 64, 3, 3)), (ops.spec_head_ps(3), (576, 64, 3, 3))):
        idx = ops.pack_index(spec)
        real = idx[idx >= 0]
        assert len(real) == len(set(real.tolist())) == int(np.prod(shape))
    idx = ops.pack_index(ops.spec_refine_conv1())
    real = idx[idx >= 0]
    assert len(set(real.tolist())) == len(real) == 129 * 640 * 9            # every non-positional input channel once
    # transposed (data-gradient) operand of the ConvLSTM conv covers the same elements
    t = ops._spec(256, 128, 3, 1, [0], 256, 4, 9, 128, transpose_flip=1)
    real = ops.pack_index(t)
    assert sorted(real[real >= 0].tolist()) == list(range(256 * 128 * 9))


def test_config_registry_resolves_reference_configs():
    import src
    from src.main import Config, _get_instance
    cfg = Config({'net': {'name': 'RefineNet', 'kwargs': dict(in_channels=1, out_channels=1, num_features=[64, 64, 64],
                                                               upscale_factor=4, num_stages=3, update_memory=True,
                                                               num_updated_frames=6, refine_window_size=5,
                                                               positional_encoding=True)},
                  'monitor': {'name': 'Monitor', 'kwargs': {'mode': 'min', 'target': 'Loss', 'saved_freq': 10}}})
    net = _get_instance(src.model.nets, cfg.net)
    assert sum(p.numel() for p in net.parameters()) == 2890993 and len(net.state_dict()) == 26
    assert 'Trainable parameters' in repr(net)
    cfg.net.kwargs.update(upscale_factor=3)
    assert cfg.net.kwargs.upscale_factor == 3 and cfg.to_dict()['net']['kwargs']['upscale_factor'] == 3
    for name in ('AcdcVSRRefineNetTrainer',):
        assert hasattr(src.runner.trainers, name)
    for name in ('AcdcVSRRefineNetPredictor',):
        assert hasattr(src.runner.predictors, name)
    for name in ('AcdcVSRRefineNetDataset', 'Dsb15VSRRefineNetDataset', 'SyntheticCineDataset'):
        assert hasattr(src.data.datasets, name)
    assert hasattr(src.data.dataloader, 'Dataloader') and hasattr(src.callbacks.loggers, 'AcdcVSRLogger')
    for name in ('PSNR', 'SSIM', 'CardiacPSNR', 'CardiacSSIM'):
        assert hasattr(src.model.metrics, name)


def test_metrics_match_reference_values():
    """PSNR / SSIM / CardiacPSNR / CardiacSSIM / denormalize against numbers produced by the reference's
    src/model/metrics.py (:9-165) and src/utils.py (oracle/make_golden.py: metrics_golden)."""
    import pickle
    import tempfile
    from src.model.metrics import PSNR, SSIM, CardiacPSNR, CardiacSSIM
    from src.utils import denormalize
    with open(os.path.join(GOLDEN, "metrics.json")) as f:
        recs = json.load(f)
    for rec in recs:
        g = torch.Generator().manual_seed(rec["seed"])
        a = torch.randn(*rec["shape"], generator=g)
        b = a + 0.1 * torch.randn(*rec["shape"], generator=g)
        a0 = a.clone()
        with tempfile.TemporaryDirectory() as td:       # metrics.py:116-165: per-patient box read from a pickle
            cp = os.path.join(td, "coordinates.pkl")
            with open(cp, "wb") as f:
                pickle.dump({"patient_g": tuple(rec["cardiac_box"])}, f)
            cpsnr, cssim = CardiacPSNR(coordinates_path=cp), CardiacSSIM(coordinates_path=cp)
            cpsnr_n = CardiacPSNR(coordinates_path=cp, size_average=False)
            cssim_n = CardiacSSIM(coordinates_path=cp, size_average=False)
        for ds in ("acdc", "dsb15"):
            da, db = denormalize(a, ds), denormalize(b, ds)
            assert abs(float(cpsnr(da, db, "patient_g")) - rec[ds]["cardiac_psnr"]) <= 1e-4
            assert abs(float(cssim(da, db, "patient_g")) - rec[ds]["cardiac_ssim"]) <= 2e-6
            assert np.allclose(cpsnr_n(da, db, "patient_g").tolist(), rec[ds]["cardiac_psnr_per_sample"], atol=1e-4)
            assert np.allclose(cssim_n(da, db, "patient_g").tolist(), rec[ds]["cardiac_ssim_per_sample"], atol=2e-6)
            assert torch.equal(a, a0)                                        # input untouched
            assert float(da.double().sum()) == rec[ds]["denorm_sum"]
            assert abs(float(PSNR()(da, db)) - rec[ds]["psnr"]) <= 1e-4       # the 0.01 dB parity bar, with margin
            assert abs(float(SSIM()(da, db)) - rec[ds]["ssim"]) <= 2e-6
            assert np.allclose(PSNR(size_average=False)(da, db).tolist(), rec[ds]["psnr_per_sample"], atol=1e-4)
            assert np.allclose(SSIM(size_average=False)(da, db).tolist(), rec[ds]["ssim_per_sample"], atol=2e-6)
    with pytest.raises(ValueError):
        denormalize(torch.zeros(1), 'other')


def test_dataset_contract_and_transforms():
    from oracle import refinenet_oracle as O
    from src.data.datasets import SyntheticCineDataset
    from src.data.datasets.acdc_vsr_refinenet_dataset import window_slices
    from src.data.transforms import compose
    T, U = 30, 6
    a, b, c, d = window_slices(T, 0, 7, U, train=False)
    frames = list(range(T))
    assert (frames * 3)[a:b] == O.circular_window(frames, T, U) and (frames * 3)[c:d] == frames
    a, b, c, d = window_slices(T, 2, 7, U, train=True)        # target frame 2: clip wraps around the cycle
    assert (frames * 3)[c:d] == [26, 27, 28, 29, 0, 1, 2] and b - a == 7 + 2 * U and a == c - U
    ds = SyntheticCineDataset(type='test', num_sequences=3)
    item = ds[2]
    assert len(item['lr_imgs']) == 42 and len(item['hr_imgs']) == 30 and item['pos_code'].shape == (42, 1)
    assert item['lr_imgs'][0].shape == (1, 54, 63) and item['hr_imgs'][0].shape == (1, 216, 252)
    assert torch.equal(item['lr_imgs'][0], item['lr_imgs'][30])               # circular padding aliases frames
    assert abs(float(item['pos_code'][6]) - 1.0) < 1e-6                       # phase 0 = end-diastole: cos(0)
    batch = torch.utils.data.default_collate([ds[0], ds[1]])
    assert batch['lr_imgs'][0].shape == (2, 1, 54, 63) and batch['pos_code'].shape == (2, 42, 1)

    np.random.seed(0)
    chain = compose([{'name': 'RandomHorizontalFlip'}, {'name': 'RandomVerticalFlip'},
                     {'name': 'RandomCropPatch', 'kwargs': {'size': [8, 8], 'ratio': 4}}])
    lr = [np.random.rand(20, 24, 1).astype(np.float32) for _ in range(3)]
    hr = [np.kron(x[..., 0], np.ones((4, 4), np.float32))[..., None] for x in lr]
    out = chain(*(lr + hr))
    assert all(o.shape == (8, 8, 1) for o in out[:3]) and all(o.shape == (32, 32, 1) for o in out[3:])
    for l, h in zip(out[:3], out[3:]):                                        # LR / HR windows stay aligned
        assert np.array_equal(np.kron(l[..., 0], np.ones((4, 4), np.float32)), h[..., 0])
    norm = compose([{'name': 'Normalize', 'kwargs': {'means': [54.089], 'stds': [48.084]}}, {'name': 'ToTensor'}])
    x = np.full((4, 5, 1), 54.089 + 48.084, np.float32)
    y = norm(x)
    assert isinstance(y, torch.Tensor) and torch.allclose(y, torch.ones(4, 5, 1), atol=1e-6)
    code = norm(np.array([1.0, -1.0], np.float32), normalize_tags=[False])
    assert code.tolist() == [1.0, -1.0]


def test_monitor_policy(tmp_path):
    from src.callbacks.monitor import Monitor
    m = Monitor(tmp_path / 'ck', 'min', 'Loss', saved_freq=10, early_stop=2)
    assert m.is_saved(10).name == 'model_10.pth' and m.is_saved(11) is None
    assert m.is_best({'Loss': 1.0}).name == 'model_best.pth' and m.best == 1.0
    assert m.is_best({'Loss': 1.5}) is None and not m.is_early_stopped()
    assert m.is_best({'Loss': 1.2}) is None and m.is_early_stopped()
    assert m.is_best({'Loss': 0.5}) is not None and m.not_improved_count == 0
    assert Monitor(tmp_path / 'ck2', 'max', 'PSNR', 1).early_stop == math.inf


def test_sharding_helpers():
    from pvsr import parallel
    for n, world in ((10, 1), (10, 4), (3, 8), (64, 8)):
        parts = [parallel.shard_indices(n, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    sizes = [42 * 54 * 63] * 5 + [42 * 108 * 126] * 3 + [20 * 54 * 63] * 4
    parts = [parallel.shard_indices(len(sizes), r, 4, sizes) for r in range(4)]
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) <= 1.35 * (sum(sizes) / 4)
    buckets = parallel.bucket_by_shape([(42, 54, 63), (42, 63, 48), (42, 54, 63)])
    assert buckets == {(42, 54, 63): [0, 2], (42, 63, 48): [1]}


def test_checkpoint_layout_roundtrip(tmp_path):
    """Trainer checkpoints keep the reference's keys and the 26-entry state_dict; the predictor reads 'net' only."""
    from helpers import build_net
    from src.callbacks.monitor import Monitor
    from src.runner.trainers.base_trainer import BaseTrainer
    kw = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=3, update_memory=True,
              num_updated_frames=6, refine_window_size=5, upscale_factor=4, positional_encoding=True)
    net = build_net(kw)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    tr = BaseTrainer(torch.device('cpu'), None, None, net, [torch.nn.L1Loss()], [1.0], [], opt, None, None,
                     Monitor(tmp_path / 'ck', 'min', 'Loss', 10), num_epochs=3)
    tr.np_random_seeds = [1, 2, 3]
    tr.epoch = 2
    tr.save(tmp_path / 'm.pth')
    ck = torch.load(tmp_path / 'm.pth', weights_only=False)
    assert set(ck) == {'net', 'optimizer', 'lr_scheduler', 'monitor', 'epoch', 'random_state', 'np_random_seeds'}
    assert len(ck['net']) == 26 and 'refine_block.prelu.weight' in ck['net']
    net2 = build_net(kw, seed=5)
    tr2 = BaseTrainer(torch.device('cpu'), None, None, net2, [torch.nn.L1Loss()], [1.0], [],
                      torch.optim.Adam(net2.parameters(), lr=1e-4), None, None,
                      Monitor(tmp_path / 'ck', 'min', 'Loss', 10), num_epochs=3)
    tr2.load(tmp_path / 'm.pth')
    assert tr2.epoch == 3 and all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))


def test_product_path_refuses_cpu_tensors():
    from helpers import build_net
    from pvsr.lib import PvsrError
    net = build_net(dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], num_stages=1, update_memory=True,
                         num_updated_frames=2, upscale_factor=2)).eval()
    with pytest.raises(PvsrError):
        with torch.no_grad():
            net([torch.zeros(1, 1, 8, 8)] * 6, torch.zeros(1, 6, 1))
    net.train()
    with pytest.raises(PvsrError):
        net([torch.zeros(1, 1, 8, 8)] * 6, torch.zeros(1, 6, 1))


def test_batched_scores_equal_per_frame_calls(tmp_path):
    """per_sample_scores (one call per loss / metric per launch) == the reference's per-frame calls
    (acdc_vsr_refinenet_predictor.py:64-75, :123-158), Cardiac* crops and unknown losses included."""
    import pickle
    from src.model.metrics import PSNR, SSIM, CardiacPSNR, CardiacSSIM
    from src.model.losses import HuberLoss
    from src.runner.predictors.base_predictor import per_sample_scores
    from src.utils import denormalize
    box = tmp_path / 'coordinates.pkl'
    with open(box, 'wb') as f:
        pickle.dump({'patient001': (3, 30, 4, 36), 'patient002': (0, 25, 10, 40)}, f)
    g = torch.Generator().manual_seed(5)
    out = torch.randn(6, 1, 40, 48, generator=g)
    tgt = out + 0.2 * torch.randn(6, 1, 40, 48, generator=g)
    od, td = denormalize(out, 'acdc'), denormalize(tgt, 'acdc')
    patients = ['patient001'] * 3 + ['patient002'] * 3
    loss_fns = [torch.nn.L1Loss(), torch.nn.MSELoss(), HuberLoss()]
    metric_fns = [PSNR(), SSIM(), CardiacPSNR(coordinates_path=box), CardiacSSIM(coordinates_path=box)]
    losses, metrics = per_sample_scores(loss_fns, metric_fns, out, tgt, od, td, patients)
    assert losses.shape == (6, 3) and metrics.shape == (6, 4)
    for i in range(6):
        for k, fn in enumerate(loss_fns):
            assert float(losses[i, k]) == pytest.approx(float(fn(out[i:i + 1], tgt[i:i + 1])), rel=1e-5)
        for k, fn in enumerate(metric_fns):
            extra = (patients[i],) if k >= 2 else ()
            assert float(metrics[i, k]) == pytest.approx(float(fn(od[i:i + 1], td[i:i + 1], *extra)), rel=1e-5)
    assert all(fn.size_average for fn in metric_fns[:2]) and metric_fns[2].inner.size_average
