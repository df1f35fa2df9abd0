"""Common parts of the predictors (reference src/runner/predictors/base_predictor.py:5-136)."""
import torch

from pvsr import parallel
from ..scores import _fusable, per_sample_scores  # noqa: F401  (re-exported for the predictors)
from ..trainers.base_trainer import to_device


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
