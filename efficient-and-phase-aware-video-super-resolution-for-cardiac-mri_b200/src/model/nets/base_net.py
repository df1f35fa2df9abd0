import torch.nn as nn


class BaseNet(nn.Module):
    """Base class of the nets: same contract as the reference's src/model/nets/base_net.py:5-13
    (an nn.Module whose repr reports the trainable parameter count)."""

    def __init__(self):
        super().__init__()

    def __repr__(self):
        n = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return super().__repr__() + f'\nTrainable parameters: {n / 1e6} M\nMemory usage: {(n * 4) / (1 << 20)} MB'
