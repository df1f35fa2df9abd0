"""Fused Adam over the flat parameter buffer (pvsr_adam_step): torch.optim.Adam.step semantics
(reference: torch.optim.Adam built by src/main.py:76 from configs/train/refine_net/exp1_x4.yaml:55-60) in ONE kernel
over all parameters, with the data-parallel gradient averaging folded in as `grad_scale`."""
import torch

from . import lib as L


class FusedAdam(torch.optim.Optimizer):
    """Drop-in for torch.optim.Adam(net.parameters(), ...) when the parameters were flattened with
    `net.engine.flatten_parameters()`.  amsgrad / maximize are not supported (the reference configs do not use
    them).  Parameters that never receive a gradient (the dead refine-block PReLU) see g = 0 and do not move,
    which equals torch skipping `grad is None`."""

    def __init__(self, params, flat_param, flat_grad, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.flat_param, self.flat_grad = flat_param, flat_grad
        self.exp_avg = torch.zeros_like(flat_param)
        self.exp_avg_sq = torch.zeros_like(flat_param)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=flat_param.device)
        self.grad_scale = 1.0
        self.engine = None

    @classmethod
    def for_net(cls, net, **kw):
        flat_p, flat_g = net.engine.flatten_parameters()
        opt = cls(net.parameters(), flat_p, flat_g, **kw)
        opt.engine = net.engine
        return opt

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        lib = L.load()
        L.check(lib.pvsr_adam_step(L.ptr(self.flat_param), L.ptr(self.flat_grad), L.ptr(self.exp_avg),
                                   L.ptr(self.exp_avg_sq), self.flat_param.numel(), g["lr"], g["betas"][0],
                                   g["betas"][1], g["eps"], g["weight_decay"], self.grad_scale,
                                   L.ptr(self.step_count), L.current_stream()), "pvsr_adam_step")
        if self.engine is not None:
            self.engine.params_changed()     # torch's version counters cannot see the kernel's in-place update
