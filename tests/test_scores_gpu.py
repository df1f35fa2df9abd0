"""Fused per-frame scores (pvsr_frame_scores): against the values the reference's own metrics.py produced
(tests/golden/metrics.json) and against the generic torch path of per_sample_scores, Cardiac* crops included."""
import json
import os
import pickle

import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu


def _scores(loss_fns, metric_fns, a, b, patients, dataset, fused):
    from src.runner.predictors.base_predictor import per_sample_scores
    from src.utils import denormalize
    return per_sample_scores(loss_fns, metric_fns, a, b, lambda: denormalize(a, dataset), lambda: denormalize(b, dataset),
                             patients, dataset=dataset if fused else None)


def test_fused_scores_match_reference_metric_values(pvsr_lib, tmp_path):
    """Fused kernel AND generic torch path against the reference's own PSNR / SSIM / CardiacPSNR / CardiacSSIM values
    (metrics.py:9-165; golden written by oracle/make_golden.py from the unmodified reference classes)."""
    from src.model.metrics import PSNR, SSIM, CardiacPSNR, CardiacSSIM
    with open(os.path.join(GOLDEN, "metrics.json")) as f:
        recs = json.load(f)
    for i, rec in enumerate(recs):
        g = torch.Generator().manual_seed(rec["seed"])
        a = torch.randn(*rec["shape"], generator=g)
        b = a + 0.1 * torch.randn(*rec["shape"], generator=g)
        box = tmp_path / f"coordinates{i}.pkl"
        with open(box, "wb") as f:
            pickle.dump({"patient_g": tuple(rec["cardiac_box"])}, f)
        fns = [PSNR().cuda(), SSIM().cuda(), CardiacPSNR(coordinates_path=box).cuda(),
               CardiacSSIM(coordinates_path=box).cuda()]
        for ds in ("acdc", "dsb15"):
            for fused in (True, False):
                losses, metrics = _scores([torch.nn.L1Loss()], fns, a.cuda(), b.cuda(), ["patient_g"] * a.shape[0], ds,
                                          fused)
                l1 = (a - b).abs().flatten(1).mean(dim=1)
                assert torch.allclose(losses[:, 0].cpu(), l1, rtol=1e-5, atol=0)
                assert torch.allclose(metrics[:, 0].cpu(), torch.tensor(rec[ds]["psnr_per_sample"]), atol=1e-4)
                assert torch.allclose(metrics[:, 1].cpu(), torch.tensor(rec[ds]["ssim_per_sample"]), atol=2e-6)
                assert torch.allclose(metrics[:, 2].cpu(), torch.tensor(rec[ds]["cardiac_psnr_per_sample"]), atol=1e-4), \
                    (rec["shape"], ds, fused)
                assert torch.allclose(metrics[:, 3].cpu(), torch.tensor(rec[ds]["cardiac_ssim_per_sample"]), atol=2e-6), \
                    (rec["shape"], ds, fused)


@pytest.mark.parametrize("shape", [(6, 1, 216, 252), (4, 1, 40, 48), (3, 1, 11, 75), (2, 1, 128, 128)])
def test_fused_scores_equal_generic_path(shape, tmp_path, pvsr_lib):
    from src.model.metrics import PSNR, SSIM, CardiacPSNR, CardiacSSIM
    n, _, H, W = shape
    box = tmp_path / 'coordinates.pkl'
    with open(box, 'wb') as f:
        h0, w0 = (H // 5 if H >= 30 else 0), (W // 7 if W >= 30 else 0)
        pickle.dump({'patient001': (0, H, 0, W),
                     'patient002': (h0, min(H, h0 + max(11, H // 2)), w0, min(W, w0 + max(13, W // 2)))}, f)
    g = torch.Generator().manual_seed(8)
    a = torch.randn(*shape, generator=g).cuda()
    b = (a + 0.3 * torch.randn(*shape, generator=g).cuda())
    patients = ['patient001' if i % 2 == 0 else 'patient002' for i in range(n)]
    patients.sort()
    loss_fns = [torch.nn.L1Loss()]
    metric_fns = [PSNR().cuda(), SSIM().cuda(), CardiacPSNR(coordinates_path=box).cuda(),
                  CardiacSSIM(coordinates_path=box).cuda()]
    lf, mf = _scores(loss_fns, metric_fns, a, b, patients, 'acdc', True)
    lg, mg = _scores(loss_fns, metric_fns, a, b, patients, 'acdc', False)
    assert lf.shape == lg.shape == (n, 1) and mf.shape == mg.shape == (n, 4)
    assert torch.allclose(lf, lg, rtol=1e-5, atol=0)
    assert torch.allclose(mf[:, [0, 2]], mg[:, [0, 2]], atol=1e-4)          # PSNR, CardiacPSNR (dB)
    assert torch.allclose(mf[:, [1, 3]], mg[:, [1, 3]], atol=3e-6)          # SSIM, CardiacSSIM


def test_unfusable_configurations_take_the_generic_path(pvsr_lib):
    from src.model.losses import HuberLoss
    from src.model.metrics import PSNR
    from src.runner.scores import _fusable
    x = torch.zeros(2, 1, 32, 32, device='cuda')
    assert _fusable([torch.nn.L1Loss()], [PSNR()], x)
    assert not _fusable([HuberLoss()], [PSNR()], x)
    assert not _fusable([torch.nn.L1Loss()], [PSNR(max_value=1)], x)
    assert not _fusable([torch.nn.L1Loss()], [PSNR()], x.cpu())
    assert not _fusable([torch.nn.L1Loss()], [PSNR()], torch.zeros(2, 1, 8, 32, device='cuda'))
