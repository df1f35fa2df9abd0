"""Host-side runtime of the RefineNet path: owns plans (one per input geometry), the packed bf16 parameter
buffers and the device workspace, and drives `pvsr_plan_forward` (include/pvsr.h).

torch is used for device memory and streams only; all arithmetic happens in libpvsr.so.
"""
import ctypes as C
import math

import torch

from . import lib as L

NUM_CLASSES = 7
CLASS_NAMES = ["in_conv", "convlstm_cell", "refine_conv1", "refine_conv2", "head_conv_ps", "head_conv_last", "misc"]


class _Plan:
    """One pvsr_plan + its device buffers (workspace, packed parameters, staging, output)."""

    def __init__(self, lib, cfg, device):
        self.lib = lib
        self.cfg = cfg
        h = C.c_void_p()
        L.check(lib.pvsr_plan_create(C.byref(cfg), C.byref(h)), "pvsr_plan_create")
        self.handle = h
        self.device = device
        self.ws_bytes = lib.pvsr_plan_workspace_bytes(h)
        self.pk_bytes = lib.pvsr_plan_packed_bytes(h)
        self.n_lists = lib.pvsr_plan_num_lists(h)
        self.launches = lib.pvsr_plan_num_launches(h)
        self.flops = lib.pvsr_plan_flops(h)
        self.workspace = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=device)
        self.packed = torch.zeros(self.pk_bytes, dtype=torch.uint8, device=device)
        self.T = cfg.n_frames - 2 * cfg.n_updated
        self.Hs, self.Ws = cfg.h * cfg.scale, cfg.w * cfg.scale
        self.lr = torch.empty(cfg.n_frames, cfg.batch, cfg.h, cfg.w, dtype=torch.float32, device=device)
        self.pos = torch.zeros(cfg.batch, cfg.n_frames, dtype=torch.float32, device=device)
        self.out = torch.empty(self.n_lists, self.T, cfg.batch, self.Hs, self.Ws, dtype=torch.float32, device=device)
        self.packed_version = None

    def class_stats(self):
        launches = (C.c_int64 * NUM_CLASSES)()
        flops = (C.c_double * NUM_CLASSES)()
        self.lib.pvsr_plan_class_stats(self.handle, launches, flops)
        return list(launches), list(flops)

    def __del__(self):
        try:
            if self.handle:
                self.lib.pvsr_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class RefineNetEngine:
    """Runs RefineNet.forward (reference refine_net.py:61-135) for a module exposing the reference's parameters."""

    def __init__(self, net):
        self.net = net
        self.plans = {}
        self.use_graph = True

    # -------------------------------------------------------------------------------------------- parameters
    def _named(self):
        n = self.net
        sd = dict(n.named_parameters())
        return sd

    def _params_struct(self):
        sd = self._named()
        P = L.NetParams()
        keep = []

        def dp(name):
            t = sd[name]
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise L.PvsrError(f"parameter {name} must be contiguous fp32")
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
        n_head = self.net.num_head_convs
        for q in range(n_head):
            P.head_w[q] = dp(f"out_block.conv{q + 1}.weight")
            P.head_b[q] = dp(f"out_block.conv{q + 1}.bias")
        return P, keep

    def _param_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.net.parameters())

    # -------------------------------------------------------------------------------------------- plans
    def plan_for(self, batch, n_frames, h, w, all_heads, device):
        key = (batch, n_frames, h, w, bool(all_heads), str(device))
        pl = self.plans.get(key)
        if pl is None:
            n = self.net
            cfg = L.NetConfig()
            cfg.batch, cfg.n_frames, cfg.n_updated, cfg.h, cfg.w = batch, n_frames, n.num_updated_frames, h, w
            cfg.scale, cfg.n_stages, cfg.window = n.upscale_factor, n.num_stages, n.refine_window_size
            cfg.n_layers = len(n.num_features)
            cfg.pos_enc, cfg.memory = int(n.positional_encoding), int(n.memory)
            cfg.all_heads, cfg.save_for_backward = int(bool(all_heads)), 0
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
        L.check(pl.lib.pvsr_plan_forward(pl.handle, C.byref(P), L.ptr(pl.packed), L.ptr(pl.lr), L.ptr(pl.pos),
                                         L.ptr(pl.out), L.ptr(pl.workspace), int(self.use_graph),
                                         L.current_stream()), "pvsr_plan_forward")
        return pl.out

    def forward(self, inputs, pos_codes, all_heads=True, clone=True):
        x0 = inputs[0]
        if not x0.is_cuda:
            raise L.PvsrError("RefineNet (B200) runs on CUDA only; there is no CPU fallback - move inputs to cuda")
        if x0.dim() != 4 or x0.shape[1] != 1:
            raise ValueError(f"expected frames of shape (N, 1, h, w), got {tuple(x0.shape)}")
        n, _, h, w = x0.shape
        pl = self.plan_for(n, len(inputs), h, w, all_heads, x0.device)
        self.stage_inputs(pl, inputs, pos_codes)
        out = self.run(pl)
        if clone:
            out = out.clone()
        return tuple([out[l, t].unsqueeze(1) for t in range(pl.T)] for l in range(pl.n_lists))

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
