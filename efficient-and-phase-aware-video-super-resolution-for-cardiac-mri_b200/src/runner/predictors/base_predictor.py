"""Common parts of the predictors (reference src/runner/predictors/base_predictor.py:5-136)."""
import torch

from pvsr import parallel
from ..trainers.base_trainer import to_device


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


def per_sample_scores(loss_fns, metric_fns, out, tgt, out_d, tgt_d, patients):
    """(losses (N, #loss), metrics (N, #metric)) of N frames: losses on the normalised frames, metrics on the
    de-normalised ones; `patients[i]` names the bounding box of frame i for the Cardiac* metrics.  The reference calls
    every loss / metric once per frame and reads each scalar back with .item()
    (acdc_vsr_refinenet_predictor.py:64-75); here a launch costs one call per loss / metric and one transfer."""
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


class BasePredictor:
    def __init__(self, device, test_dataloader, net, loss_fns, loss_weights, metric_fns):
        self.device = device
        self.test_dataloader = test_dataloader
        self.net = net.to(device)
        self.loss_fns = [fn.to(device) for fn in loss_fns]
        self.loss_weights = torch.tensor(loss_weights, dtype=torch.float, device=device)
        self.metric_fns = [fn.to(device) for fn in metric_fns]
        self.rank, self.world = parallel.rank_world()

    def predict(self):
        raise NotImplementedError

    def _allocate_data(self, batch):
        return to_device(batch, self.device)

    def _init_log(self):
        names = ['Loss'] + [fn.__class__.__name__ for fn in self.loss_fns + self.metric_fns]
        return dict.fromkeys(names, 0)

    def load(self, path):
        """Restores the network weights from a trainer checkpoint (only the 'net' entry is read)."""
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        self.net.load_state_dict(ckpt['net'])
        if hasattr(self.net, 'engine'):
            self.net.engine.params_changed()
