"""Per-frame losses / metrics of the runners, scored a whole launch at a time.

The reference calls every loss and metric once per frame and reads each scalar back with .item()
(src/runner/predictors/acdc_vsr_refinenet_predictor.py:64-75,123-158; trainers likewise).  `per_sample_scores` gives
the same numbers for N frames with one call per loss / metric (generic path, any torch loss / metric module) or - when
only L1Loss, PSNR, SSIM and their Cardiac* crops are configured - with the fused kernels behind pvsr_frame_scores.
"""
import torch


def _per_sample_loss(fn, out, tgt):
    """`fn(out[i:i+1], tgt[i:i+1])` for every sample i in ONE pass when fn is a mean-reduced element-wise torch loss."""
    if getattr(fn, 'reduction', None) == 'mean':
        if type(fn) is torch.nn.L1Loss:
            return (out - tgt).abs().flatten(1).mean(dim=1)
        if type(fn) is torch.nn.MSELoss:
            return (out - tgt).pow(2).flatten(1).mean(dim=1)
    return torch.stack([fn(out[i:i + 1], tgt[i:i + 1]) for i in range(out.shape[0])])


def _per_sample_metric(fn, out, tgt, *extra, chunk=256):
    """Per-sample scores of a metric module: PSNR / SSIM expose `size_average` (reference metrics.py:26,92), which is
    switched off for the call so a whole launch is scored at once; anything else falls back to one call per sample."""
    inner = getattr(fn, 'inner', fn)                       # Cardiac* wrap a PSNR / SSIM
    if hasattr(inner, 'size_average'):
        keep, inner.size_average = inner.size_average, False
        try:
            return torch.cat([fn(out[i:i + chunk], tgt[i:i + chunk], *extra) for i in range(0, out.shape[0], chunk)])
        finally:
            inner.size_average = keep
    return torch.stack([fn(out[i:i + 1], tgt[i:i + 1], *extra) for i in range(out.shape[0])])


def _fusable(loss_fns, metric_fns, out):
    """True when every configured loss / metric is one the fused kernel computes (pvsr_frame_scores): L1Loss(mean),
    PSNR(max 255), 2-D single-channel SSIM(range 255) and their Cardiac* crops, on single-channel CUDA fp32 frames."""
    from src.model.metrics import PSNR, SSIM
    if not (out.is_cuda and out.dtype == torch.float32 and out.dim() == 4 and out.shape[1] == 1):
        return False
    if out.shape[2] < 11 or out.shape[3] < 11:
        return False
    for fn in loss_fns:
        if type(fn) is not torch.nn.L1Loss or fn.reduction != 'mean':
            return False
    for fn in metric_fns:
        inner = getattr(fn, 'inner', fn)
        if type(inner) is PSNR:
            if inner.max_value != 255:
                return False
        elif type(inner) is SSIM:
            if inner.dim != 2 or inner.channels != 1 or inner.value_range != 255:
                return False
        else:
            return False
    return True


def _fused_scores(loss_fns, metric_fns, out, tgt, dataset, patients):
    """L1 / PSNR / SSIM (/ Cardiac*) of N frames through pvsr_frame_scores: two launches for the whole frames plus two
    for the cardiac boxes, instead of ~40 torch kernels per metric call."""
    from pvsr import lib as L
    from src.model.metrics import PSNR, SSIM
    from src.utils import _STATS
    mean, std = _STATS[dataset]
    lib = L.load()
    n, _, H, W = out.shape
    a, b = out.contiguous(), tgt.contiguous()
    window = next((getattr(fn, 'inner', fn).window for fn in metric_fns if type(getattr(fn, 'inner', fn)) is SSIM), None)
    if window is None:
        window = SSIM().window.to(out.device)
    window = window.to(device=out.device, dtype=torch.float32).contiguous()

    def run(rects):
        sums = torch.empty(n, 3, dtype=torch.float64, device=out.device)
        L.check(lib.pvsr_frame_scores(L.ptr(a), L.ptr(b), L.ptr(rects), n, H, W, mean, std, L.ptr(window), 255.0,
                                      L.ptr(sums), L.current_stream()), 'pvsr_frame_scores')
        if rects is None:
            area = torch.full((n,), float(H * W), dtype=torch.float64, device=out.device)
            valid = torch.full((n,), float((H - 10) * (W - 10)), dtype=torch.float64, device=out.device)
        else:
            r = rects.double()
            area = (r[:, 1] - r[:, 0]) * (r[:, 3] - r[:, 2])
            valid = (r[:, 1] - r[:, 0] - 10) * (r[:, 3] - r[:, 2] - 10)
        l1 = sums[:, 0] / area
        psnr = 10 * torch.log10(255.0 ** 2 / (sums[:, 1] / area + 1e-10))
        ssim = sums[:, 2] / valid
        return l1.float(), psnr.float(), ssim.float()

    full = run(None)
    crop = None
    if any(hasattr(fn, 'inner') for fn in metric_fns):
        coords = next(fn.coordinates for fn in metric_fns if hasattr(fn, 'inner'))
        boxes = [coords[p] for p in patients]
        if any(hn - h0 < 11 or wn - w0 < 11 for h0, hn, w0, wn in boxes):
            raise ValueError('a cardiac bounding box is smaller than the 11x11 SSIM window')
        rects = torch.tensor(boxes, dtype=torch.int32, device=out.device)
        crop = run(rects)
    losses = [full[0] for _ in loss_fns]
    metrics = []
    for fn in metric_fns:
        src = crop if hasattr(fn, 'inner') else full
        metrics.append(src[1] if type(getattr(fn, 'inner', fn)) is PSNR else src[2])
    stack = lambda cols: torch.stack(cols, dim=1) if cols else out.new_zeros(n, 0)
    return stack(losses), stack(metrics)


def per_sample_scores(loss_fns, metric_fns, out, tgt, out_d, tgt_d, patients, dataset=None):
    """(losses (N, #loss), metrics (N, #metric)) of N frames: losses on the normalised frames, metrics on the
    de-normalised ones; `patients[i]` names the bounding box of frame i for the Cardiac* metrics.  The reference calls
    every loss / metric once per frame and reads each scalar back with .item()
    (acdc_vsr_refinenet_predictor.py:64-75); here a launch costs one call per loss / metric and one transfer - or,
    with `dataset` given and only L1 / PSNR / SSIM / Cardiac* configured, two fused kernels (pvsr_frame_scores).
    `out_d` / `tgt_d` may be callables returning the de-normalised frames (evaluated only on the generic path)."""
    if dataset is not None and _fusable(loss_fns, metric_fns, out):
        cardiac = [fn for fn in metric_fns if hasattr(fn, 'inner')]
        if all(fn.coordinates is cardiac[0].coordinates or fn.coordinates == cardiac[0].coordinates for fn in cardiac):
            return _fused_scores(loss_fns, metric_fns, out, tgt, dataset, patients)
    out_d = out_d() if callable(out_d) else out_d
    tgt_d = tgt_d() if callable(tgt_d) else tgt_d
    losses = [_per_sample_loss(fn, out, tgt).float() for fn in loss_fns]
    metrics = []
    for fn in metric_fns:
        if 'Cardiac' in fn.__class__.__name__:
            parts, i = [], 0
            while i < len(patients):                           # runs of frames of one patient share a box
                j = i
                while j < len(patients) and patients[j] == patients[i]:
                    j += 1
                parts.append(_per_sample_metric(fn, out_d[i:j], tgt_d[i:j], patients[i]))
                i = j
            metrics.append(torch.cat(parts).float())
        else:
            metrics.append(_per_sample_metric(fn, out_d, tgt_d).float())
    n = out.shape[0]
    stack = lambda cols: torch.stack(cols, dim=1) if cols else out.new_zeros(n, 0)
    return stack(losses), stack(metrics)
