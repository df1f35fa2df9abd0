"""Checkpoint policy (reference src/callbacks/monitor.py:4-63): a periodic checkpoint every `saved_freq` epochs,
`model_best.pth` whenever the monitored validation value improves, optional early stopping.  The object is pickled
into checkpoints by the trainer (base_trainer.py:233), so its attribute names follow the reference."""
import math
from pathlib import Path


class Monitor:
    def __init__(self, checkpoints_dir, mode, target, saved_freq, early_stop=0):
        if mode not in ('min', 'max'):
            raise ValueError(f"The monitor mode should be 'min' or 'max'. Got {mode}.")
        self.checkpoints_dir = Path(checkpoints_dir)
        self.mode, self.target, self.saved_freq = mode, target, saved_freq
        self.early_stop = early_stop if early_stop else math.inf
        self.best = math.inf if mode == 'min' else -math.inf
        self.not_improved_count = 0
        self.checkpoints_dir.mkdir(parents=True, exist_ok=True)

    def is_saved(self, epoch):
        """Path of the periodic checkpoint due at `epoch`, else None."""
        return self.checkpoints_dir / f'model_{epoch}.pth' if epoch % self.saved_freq == 0 else None

    def is_best(self, valid_log):
        """Path of the best-checkpoint file if `valid_log[target]` improved (and records it), else None."""
        score = valid_log[self.target]
        improved = score < self.best if self.mode == 'min' else score > self.best
        if not improved:
            self.not_improved_count += 1
            return None
        self.best, self.not_improved_count = score, 0
        return self.checkpoints_dir / 'model_best.pth'

    def is_early_stopped(self):
        return self.not_improved_count == self.early_stop
