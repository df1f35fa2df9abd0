"""Fused Adam over the flat parameter buffer (pvsr_adam_step): torch.optim.Adam.step semantics
(reference: torch.optim.Adam built by src/main.py:76 from configs/train/refine_net/exp1_x4.yaml:55-60) in ONE kernel
over all parameters, with the data-parallel gradient averaging folded in as `grad_scale`."""
import torch

from . import lib as L


class FusedAdam(torch.optim.Optimizer):
    """Drop-in for torch.optim.Adam(net.parameters(), ...) when the parameters were flattened with
    `net.engine.flatten_parameters()`.  amsgrad / maximize are not supported (the reference configs do not use
    them).  Parameters that never receive a gradient (the dead refine-block PReLU, `engine.dead_parameters`) see
    g = 0 and do not move as long as weight_decay == 0 (the reference configs); with weight decay their slots are put
    back after the kernel (value and moments), which equals torch skipping `grad is None`."""

    def __init__(self, params, flat_param, flat_grad, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.flat_param, self.flat_grad = flat_param, flat_grad
        self.exp_avg = torch.zeros_like(flat_param)
        self.exp_avg_sq = torch.zeros_like(flat_param)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=flat_param.device)
        self.grad_scale = 1.0
        self.engine = None
        self.frozen = []          # (offset, numel) of parameters that never receive a gradient (see for_net)

    @classmethod
    def for_net(cls, net, **kw):
        flat_p, flat_g = net.engine.flatten_parameters()
        opt = cls(net.parameters(), flat_p, flat_g, **kw)
        opt.engine = net.engine
        dead = set(getattr(net.engine, 'dead_parameters', ()))
        named = dict(net.named_parameters())
        es, base = flat_p.element_size(), flat_p.data_ptr()
        opt.frozen = [((named[k].data_ptr() - base) // es, named[k].numel()) for k in dead if k in named]
        return opt

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()

    # ---- checkpointing: torch.optim.Adam's layout, so FusedAdam and torch Adam checkpoints interchange
    def _slices(self):
        """(parameter, element offset in the flat buffer) in param_groups order."""
        base, es = self.flat_param.data_ptr(), self.flat_param.element_size()
        out = []
        for g in self.param_groups:
            for p in g["params"]:
                off = (p.data_ptr() - base) // es
                if off < 0 or off + p.numel() > self.flat_param.numel():
                    raise L.PvsrError("FusedAdam: a parameter is not a view of the flat buffer")
                out.append((p, off))
        return out

    def state_dict(self):
        """{'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]} exactly like torch.optim.Adam
        (base_trainer.py:230 saves it, :245 loads it); moments are split per parameter out of the flat buffers.
        Parameters without a gradient so far (step 0) carry no state, as in torch."""
        sd = super().state_dict()
        steps = float(self.step_count.item())
        state = {}
        if steps > 0:
            for i, (p, off) in enumerate(self._slices()):
                n = p.numel()
                state[i] = {"step": torch.tensor(steps, dtype=torch.float32),
                            "exp_avg": self.exp_avg[off:off + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).clone()}
        sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        """Accepts FusedAdam's own checkpoints and torch.optim.Adam's (same keys).  Parameters missing from `state`
        (torch skips parameters that never received a gradient, e.g. the dead refine-block PReLU) restart from zero
        moments; the step count is the maximum over the stored per-parameter steps (identical for every parameter that
        has state in both optimisers)."""
        state = state_dict.get("state", {})
        groups = state_dict.get("param_groups", [])
        for g, sg in zip(self.param_groups, groups):
            for k, v in sg.items():
                if k != "params":
                    g[k] = v
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = 0.0
        with torch.no_grad():
            for i, (p, off) in enumerate(self._slices()):
                st = state.get(i, state.get(str(i)))
                if not st:
                    continue
                n = p.numel()
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1).to(self.exp_avg.device, torch.float32))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1).to(self.exp_avg.device, torch.float32))
                steps = max(steps, float(st["step"]))
            self.step_count.fill_(steps)

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        lib = L.load()
        keep = [(o, n, self.flat_param[o:o + n].clone()) for o, n in self.frozen] if g["weight_decay"] != 0 else []
        self._adam_kernel(lib, g)
        for o, n, v in keep:      # torch.optim.Adam skips parameters whose grad is None: no decay, no state
            self.flat_param[o:o + n].copy_(v)
            self.exp_avg[o:o + n].zero_()
            self.exp_avg_sq[o:o + n].zero_()
        if self.engine is not None:
            self.engine.params_changed()     # torch's version counters cannot see the kernel's in-place update

    def _adam_kernel(self, lib, g):
        L.check(lib.pvsr_adam_step(L.ptr(self.flat_param), L.ptr(self.flat_grad), L.ptr(self.exp_avg),
                                   L.ptr(self.exp_avg_sq), self.flat_param.numel(), g["lr"], g["betas"][0],
                                   g["betas"][1], g["eps"], g["weight_decay"], self.grad_scale,
                                   L.ptr(self.step_count), L.current_stream()), "pvsr_adam_step")
