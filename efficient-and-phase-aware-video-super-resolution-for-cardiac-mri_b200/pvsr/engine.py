"""Host-side runtime of the RefineNet path: owns plans (one per input geometry), the packed bf16 parameter
buffers and the device workspace, and drives `pvsr_plan_forward` / `pvsr_plan_backward` (include/pvsr.h).

torch is used for device memory and streams only; all arithmetic happens in libpvsr.so.
"""
import ctypes as C

import torch

from . import lib as L

NUM_CLASSES = L.NUM_CLASSES
CLASS_NAMES = ["in_conv", "convlstm_cell", "refine_conv1", "refine_conv2", "head_conv_ps", "head_conv_last", "misc"]
CLASS_NAMES_BWD = ["head_last_bwd", "head_dgrad", "head_wgrad", "refine_dgrad", "refine_wgrad", "lstm_pointwise_bwd",
                   "lstm_dgrad", "lstm_wgrad", "misc_bwd"]


class _Plan:
    """One pvsr_plan + its device buffers (workspace, packed parameters, staging, output)."""

    def __init__(self, lib, cfg, device):
        self.lib = lib
        self.cfg = cfg
        h = C.c_void_p()
        L.check(lib.pvsr_plan_create(C.byref(cfg), C.byref(h)), "pvsr_plan_create")
        self.handle = h
        self.device = device
        self.train = bool(cfg.save_for_backward)
        self.ws_bytes = lib.pvsr_plan_workspace_bytes(h)
        self.pk_bytes = lib.pvsr_plan_packed_bytes(h)
        self.n_lists = lib.pvsr_plan_num_lists(h)
        self.launches = lib.pvsr_plan_num_launches(h)
        self.flops = lib.pvsr_plan_flops(h)
        self.launches_bwd = lib.pvsr_plan_num_launches_bwd(h)
        self.flops_bwd = lib.pvsr_plan_flops_bwd(h)
        self.workspace = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=device)
        self.packed = torch.zeros(self.pk_bytes, dtype=torch.uint8, device=device)
        self.T = cfg.n_frames - 2 * cfg.n_updated
        self.Hs, self.Ws = cfg.h * cfg.scale, cfg.w * cfg.scale
        self.lr = torch.empty(cfg.n_frames, cfg.batch, cfg.h, cfg.w, dtype=torch.float32, device=device)
        self.pos = torch.zeros(cfg.batch, cfg.n_frames, dtype=torch.float32, device=device)
        self.out = torch.empty(self.n_lists, self.T, cfg.batch, self.Hs, self.Ws, dtype=torch.float32, device=device)
        self.dout = torch.zeros_like(self.out) if self.train else None
        self.target = torch.empty(self.T, cfg.batch, self.Hs, self.Ws, dtype=torch.float32,
                                  device=device) if self.train else None
        self.packed_version = None
        self.outs, self.out_idx = [self.out], 0     # output ring (RefineNetEngine.output_slots)

    def class_stats(self, bwd=False):
        n = L.NUM_CLASSES_BWD if bwd else NUM_CLASSES
        launches = (C.c_int64 * n)()
        flops = (C.c_double * n)()
        (self.lib.pvsr_plan_class_stats_bwd if bwd else self.lib.pvsr_plan_class_stats)(self.handle, launches, flops)
        return list(launches), list(flops)

    def __del__(self):
        try:
            if self.handle:
                self.lib.pvsr_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def bind_reference_module(net):
    """RefineNetEngine for an instance of the REFERENCE's own `RefineNet` (src/model/nets/refine_net.py:10-135), i.e.
    a module this package did not construct.  The engine reads seven scalars and
    `named_parameters()`; the reference stores four of them on the module (:21-28) and the other three are derived
    from its sub-modules here: `memory` (ConvLSTMCell.memory, :231), `positional_encoding` (_RefineBlock, :145) and
    the number of convolutions in `_OutBlock` (:194-205)."""
    if not hasattr(net, 'memory'):
        net.memory = bool(net.forward_lstm_block.cell_list[0].memory)
    if not hasattr(net, 'positional_encoding'):
        net.positional_encoding = bool(net.refine_block.positional_encoding)
    if not hasattr(net, 'num_head_convs'):
        net.num_head_convs = sum(1 for m in net.out_block.children() if isinstance(m, torch.nn.Conv2d))
    if net.in_channels != 1 or net.out_channels != 1 or any(f != 64 for f in net.num_features):
        raise ValueError('The B200 path implements the 1 -> 64 -> 1 channel configuration of the reference configs '
                         f'(got in/out channels {net.in_channels}/{net.out_channels}, features {net.num_features}).')
    engine = RefineNetEngine(net)
    net.engine = engine          # pvsr.autograd.refinenet_train_forward looks the engine up on the module
    return engine


class RefineNetEngine:
    """Runs RefineNet.forward / backward (reference refine_net.py:61-135) for a module exposing the reference's
    parameters."""

    # parameters the reference registers but never reads (refine_net.py:147 builds `_RefineBlock.prelu`, forward :157-185
    # never applies it): their grad stays None in torch, so optimisers must leave them alone (pvsr.optim.FusedAdam)
    dead_parameters = ('refine_block.prelu.weight',)

    def __init__(self, net):
        self.net = net
        self.plans = {}
        self.use_graph = True
        # inference plans rotate over this many output buffers: with 2, the device->host copy of step i (on a copy
        # stream, pvsr.hostio.HostFrameRing) overlaps the whole forward of step i + 1 instead of blocking it
        self.output_slots = 1
        self._flat = None   # (flat_param, flat_grad, views) once flatten_parameters() ran
        self._grad_buf = None
        self._aux = {}

    # -------------------------------------------------------------------------------------------- parameters
    def _named(self):
        return dict(self.net.named_parameters())

    def _fill_struct(self, P, tensors, keep):
        """Fills a NetParams / NetGrads structure from a {name: tensor} mapping (missing names stay NULL)."""

        def dp(name):
            t = tensors.get(name)
            if t is None:
                return None
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise L.PvsrError(f"{name} must be contiguous fp32")
            keep.append(t)
            return t.data_ptr()

        P.in_w, P.in_b, P.in_slope = dp("in_block.conv.weight"), dp("in_block.conv.bias"), dp("in_block.prelu.weight")
        nl = len(self.net.num_features)
        for d, pre in enumerate(("forward_lstm_block", "backward_lstm_block")):
            for l in range(nl):
                P.lstm_w[d][l] = dp(f"{pre}.cell_list.{l}.conv.weight")
                P.lstm_b[d][l] = dp(f"{pre}.cell_list.{l}.conv.bias")
        P.ref_w1, P.ref_b1 = dp("refine_block.body.conv1.weight"), dp("refine_block.body.conv1.bias")
        if self.net.positional_encoding:
            P.ref_w2, P.ref_b2 = dp("refine_block.body.conv2.weight"), dp("refine_block.body.conv2.bias")
        for q in range(self.net.num_head_convs):
            P.head_w[q] = dp(f"out_block.conv{q + 1}.weight")
            P.head_b[q] = dp(f"out_block.conv{q + 1}.bias")
        return P

    def _params_struct(self):
        keep = []
        P = self._fill_struct(L.NetParams(), self._named(), keep)
        return P, keep

    def _param_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.net.parameters())

    def params_changed(self):
        """Marks the packed bf16 operands stale (for in-place parameter updates torch's version counter misses)."""
        for pl in self.plans.values():
            pl.packed_version = None

    def flatten_parameters(self):
        """Re-points every parameter (and its .grad) at a slice of ONE flat fp32 buffer, so that the data-parallel
        gradient exchange is a single NCCL all-reduce and the optimiser a single kernel (pvsr.optim.FusedAdam).
        Values are preserved; call after .to(device) and before constructing the optimiser state."""
        if self._flat is not None:
            return self._flat
        params = list(self.net.parameters())
        dev = params[0].device
        sizes = [p.numel() for p in params]
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += (n + 3) // 4 * 4          # keep every slice 16-byte aligned
        flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o, n in zip(params, offs, sizes):
                flat_p[o:o + n].copy_(p.detach().reshape(-1))
                p.data = flat_p[o:o + n].view(p.shape)
                p.grad = flat_g[o:o + n].view(p.shape)
        self._flat = (flat_p, flat_g)
        self.params_changed()
        return self._flat

    # -------------------------------------------------------------------------------------------- plans
    def plan_for(self, batch, n_frames, h, w, all_heads, device, train=False):
        key = (batch, n_frames, h, w, bool(all_heads), str(device), bool(train))
        pl = self.plans.get(key)
        if pl is None:
            n = self.net
            cfg = L.NetConfig()
            cfg.batch, cfg.n_frames, cfg.n_updated, cfg.h, cfg.w = batch, n_frames, n.num_updated_frames, h, w
            cfg.scale, cfg.n_stages, cfg.window = n.upscale_factor, n.num_stages, n.refine_window_size
            cfg.n_layers = len(n.num_features)
            cfg.pos_enc, cfg.memory = int(n.positional_encoding), int(n.memory)
            cfg.all_heads, cfg.save_for_backward = int(bool(all_heads)), int(bool(train))
            pl = _Plan(L.load(), cfg, device)
            self.plans[key] = pl
        return pl

    def _ensure_packed(self, pl, P):
        ver = self._param_version()
        if pl.packed_version != ver:
            L.check(pl.lib.pvsr_plan_pack(pl.handle, C.byref(P), L.ptr(pl.packed), L.current_stream()),
                    "pvsr_plan_pack")
            pl.packed_version = ver

    # -------------------------------------------------------------------------------------------- forward
    def stage_inputs(self, pl, inputs, pos_codes):
        torch.stack([x.reshape(pl.cfg.batch, pl.cfg.h, pl.cfg.w) for x in inputs], dim=0, out=pl.lr)
        if pos_codes is not None:
            pl.pos.copy_(pos_codes.reshape(pl.cfg.batch, pl.cfg.n_frames))

    def run(self, pl):
        """Enqueues one forward over the staged inputs; returns the plan's output buffer
        [lists, T, N, H*s, W*s] (reused by the next call)."""
        P, keep = self._params_struct()
        self._ensure_packed(pl, P)
        pl.out = self._next_out(pl, advance=True)
        L.check(pl.lib.pvsr_plan_forward(pl.handle, C.byref(P), L.ptr(pl.packed), L.ptr(pl.lr), L.ptr(pl.pos),
                                         L.ptr(pl.out), L.ptr(pl.workspace), int(self.use_graph),
                                         L.current_stream()), "pvsr_plan_forward")
        return pl.out

    def _next_out(self, pl, advance=False):
        slots = 1 if pl.train else max(1, int(self.output_slots))
        while len(pl.outs) < slots:
            pl.outs.append(torch.empty_like(pl.outs[0]))
        buf = pl.outs[pl.out_idx % slots]
        if advance:
            pl.out_idx += 1
        return buf

    def next_output_ptr(self, pl):
        """Device address of the buffer the next run(pl) writes (for HostFrameRing.before_launch)."""
        return self._next_out(pl).data_ptr()

    def next_output_bytes(self, pl):
        buf = self._next_out(pl)
        return buf.numel() * buf.element_size()

    def _check_inputs(self, inputs):
        x0 = inputs[0]
        if not x0.is_cuda:
            raise L.PvsrError("RefineNet (B200) runs on CUDA only; there is no CPU fallback - move inputs to cuda")
        if x0.dim() != 4 or x0.shape[1] != 1:
            raise ValueError(f"expected frames of shape (N, 1, h, w), got {tuple(x0.shape)}")
        return x0

    def forward(self, inputs, pos_codes, all_heads=True, clone=True):
        x0 = self._check_inputs(inputs)
        n, _, h, w = x0.shape
        pl = self.plan_for(n, len(inputs), h, w, all_heads, x0.device)
        self.stage_inputs(pl, inputs, pos_codes)
        out = self.run(pl)
        if clone:
            out = out.clone()
        return tuple([out[l, t].unsqueeze(1) for t in range(pl.T)] for l in range(pl.n_lists))

    # -------------------------------------------------------------------------------------------- training
    def train_plan(self, inputs):
        x0 = self._check_inputs(inputs)
        n, _, h, w = x0.shape
        return self.plan_for(n, len(inputs), h, w, True, x0.device, train=True)

    def backward(self, pl, grads, generic=False):
        """Enqueues the backward pass of the last `run(pl)`; pl.dout holds dL/d(out).  `grads`: {parameter name:
        fp32 tensor} accumulated in place (+=).  generic=True: pl.dout is an arbitrary gradient (autograd path)."""
        if generic:
            L.check(pl.lib.pvsr_plan_set_sign_gradient(pl.handle, None, 0), "pvsr_plan_set_sign_gradient")
        P, keep = self._params_struct()
        G = self._fill_struct(L.NetGrads(), grads, keep)
        L.check(pl.lib.pvsr_plan_backward(pl.handle, C.byref(P), L.ptr(pl.packed), L.ptr(pl.lr), L.ptr(pl.pos),
                                          L.ptr(pl.dout), C.byref(G), L.ptr(pl.workspace), int(self.use_graph),
                                          L.current_stream()), "pvsr_plan_backward")

    def grad_buffers(self):
        """Per-parameter fp32 gradient buffers reused by every autograd backward (zeroed by the caller)."""
        if self._grad_buf is None:
            self._grad_buf = {k: torch.zeros_like(p) for k, p in self._named().items()}
        return self._grad_buf

    def loss_and_grads(self, inputs, pos_codes, targets, loss_weights=None, zero_grads=True):
        """Fused training step body: forward, the trainer's multi-stage L1 loss
        (acdc_vsr_refinenet_trainer.py:83-93) and backward, without autograd.  Gradients are accumulated into
        `p.grad` (allocated / zeroed here).  Returns (loss: 0-dim tensor, out buffer [lists, T, N, Hs, Ws])."""
        pl = self.train_plan(inputs)
        self.stage_inputs(pl, inputs, pos_codes)
        torch.stack([t.reshape(pl.cfg.batch, pl.Hs, pl.Ws) for t in targets], dim=0, out=pl.target)
        out = self.run(pl)
        S = pl.n_lists // 3
        n_per_list = pl.T * pl.cfg.batch * pl.Hs * pl.Ws
        sign_scales = None
        if loss_weights is None:
            key = ("lw", pl.n_lists, n_per_list, str(pl.device))
            loss_weights = self._aux.get(key)
            if loss_weights is None:
                w = [0.5 ** (S - k // 3 - 1) / n_per_list for k in range(pl.n_lists)]
                loss_weights = torch.tensor(w, dtype=torch.float32, device=pl.device)
                self._aux[key] = loss_weights
                self._aux[key + ("host",)] = (C.c_float * pl.n_lists)(*loss_weights.cpu().tolist())
            sign_scales = self._aux[key + ("host",)]
        # dout = w_k * sign(out - target): tell the plan, so the tail adjoint can use the exact sign (include/pvsr.h)
        L.check(pl.lib.pvsr_plan_set_sign_gradient(pl.handle, sign_scales, pl.n_lists if sign_scales is not None else 0),
                "pvsr_plan_set_sign_gradient")
        loss = torch.zeros((), dtype=torch.float32, device=pl.device)
        L.check(pl.lib.pvsr_l1_multistage(L.ptr(out), L.ptr(pl.target), L.ptr(loss_weights), pl.n_lists, n_per_list,
                                          L.ptr(loss), L.ptr(pl.dout), L.current_stream()), "pvsr_l1_multistage")
        grads = {}
        for k, p in self._named().items():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads[k] = p.grad
        if zero_grads:
            if self._flat is not None:
                self._flat[1].zero_()
            else:
                for g in grads.values():
                    g.zero_()
        self.backward(pl, grads)
        return loss, out

    # -------------------------------------------------------------------------------------------- profiling
    def profile(self, pl):
        """One eager forward with per-launch CUDA events: {class name: (ms, launches, flops)}."""
        P, keep = self._params_struct()
        self._ensure_packed(pl, P)
        ms = (C.c_double * NUM_CLASSES)()
        L.check(pl.lib.pvsr_plan_profile(pl.handle, C.byref(P), L.ptr(pl.packed), L.ptr(pl.lr), L.ptr(pl.pos),
                                         L.ptr(pl.out), L.ptr(pl.workspace), ms, L.current_stream()),
                "pvsr_plan_profile")
        launches, flops = pl.class_stats()
        return {CLASS_NAMES[i]: (ms[i], launches[i], flops[i]) for i in range(NUM_CLASSES)}

    def profile_backward(self, pl):
        """One eager backward (of the last forward; pl.dout as is) with per-launch CUDA events."""
        P, keep = self._params_struct()
        grads = {k: torch.zeros_like(p) for k, p in self._named().items()}
        G = self._fill_struct(L.NetGrads(), grads, keep)
        ms = (C.c_double * L.NUM_CLASSES_BWD)()
        L.check(pl.lib.pvsr_plan_profile_bwd(pl.handle, C.byref(P), L.ptr(pl.packed), L.ptr(pl.lr), L.ptr(pl.pos),
                                             L.ptr(pl.dout), C.byref(G), L.ptr(pl.workspace), ms, L.current_stream()),
                "pvsr_plan_profile_bwd")
        launches, flops = pl.class_stats(bwd=True)
        return {CLASS_NAMES_BWD[i]: (ms[i], launches[i], flops[i]) for i in range(L.NUM_CLASSES_BWD)}
