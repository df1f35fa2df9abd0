"""Bicubic - the interpolation baseline of the comparison table (reference src/model/nets/bicubic.py:8-19:
`nn.Upsample(scale_factor=upscale_factor, mode='bicubic', align_corners=True)`), as one HBM-bound CUDA kernel
(pvsr_bicubic_upsample).  No parameters; `src.main` skips checkpoint loading for it (src/main.py:182)."""
import torch

from .base_net import BaseNet


class Bicubic(BaseNet):
    def __init__(self, upscale_factor):
        super().__init__()
        if int(upscale_factor) != upscale_factor or upscale_factor < 1:
            raise ValueError(f'The B200 path needs an integer upscale factor. Got {upscale_factor}.')
        self.upscale_factor = int(upscale_factor)

    def forward(self, input):
        from pvsr import lib as L
        if not input.is_cuda:
            raise L.PvsrError('Bicubic (B200) runs on CUDA only; there is no CPU fallback - move inputs to cuda')
        n, c, h, w = input.shape
        s = self.upscale_factor
        x = input.detach().contiguous().float()
        out = torch.empty(n, c, h * s, w * s, dtype=torch.float32, device=input.device)
        L.check(L.load().pvsr_bicubic_upsample(L.ptr(x), L.ptr(out), n * c, h, w, s, L.current_stream()),
                'pvsr_bicubic_upsample')
        return out
