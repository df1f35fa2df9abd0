"""Epoch logger (reference src/callbacks/loggers/base_logger.py:5-59): train/valid scalars per epoch plus an image
panel.  TensorBoard is used when importable; otherwise scalars go to `<log_dir>/scalars.jsonl` and images to
`<log_dir>/<tag>_<epoch>.pt`, so training runs on boxes without tensorboard (this image has none)."""
import json
from pathlib import Path

import torch


class BaseLogger:
    def __init__(self, log_dir, net=None, dummy_input=None):
        self.log_dir = Path(log_dir)
        self.log_dir.mkdir(parents=True, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(str(self.log_dir))
        except Exception:           # tensorboard not installed
            self.writer = None
            self._scalars = open(self.log_dir / 'scalars.jsonl', 'a')

    def write(self, epoch, train_log, train_batch, train_outputs, valid_log, valid_batch, valid_outputs):
        self._add_scalars(epoch, train_log, valid_log)
        self._add_images(epoch, train_batch, train_outputs, valid_batch, valid_outputs)

    def close(self):
        if self.writer is not None:
            self.writer.close()
        else:
            self._scalars.close()

    def _add_scalars(self, epoch, train_log, valid_log):
        if self.writer is not None:
            for key in train_log:
                self.writer.add_scalars(key, {'train': train_log[key], 'valid': valid_log[key]}, epoch)
        else:
            self._scalars.write(json.dumps({'epoch': epoch, 'train': train_log, 'valid': valid_log}) + '\n')
            self._scalars.flush()

    def _add_image(self, tag, img, epoch):
        if self.writer is not None:
            self.writer.add_image(tag, img)
        else:
            torch.save(img.cpu(), self.log_dir / f'{tag}_{epoch}.pt')

    def _add_images(self, epoch, train_batch, train_outputs, valid_batch, valid_outputs):
        raise NotImplementedError
