"""Host-side runtime of DRFNet (reference src/model/nets/drf_net.py:8-147) on the RefineNet conv
core: every convolution of the net - including the projection units' ConvTranspose2d / strided Conv2d - and every
data / weight gradient is a launch of the tcgen05 implicit-GEMM kernels behind `pvsr_conv3x3_fwd` /
`pvsr_conv3x3_wgrad_multi` (include/pvsr.h).

The feedback block never leaves the LR grid.  With s = upscale_factor, P = s*s and (k, s, p) the projection geometry
(drf_net.py:69-76: k = s + 4, p = 2), an HR feature map is kept PHASE-STACKED: [n, h, w, P*F] with channel q*F + c =
HR pixel (s*y + q // s, s*x + q % s), channel c.  Then

    deconv (ConvTranspose2d k, s, p)   = 3x3 conv  F -> P*F   tap (dy, dx), phase (ry, rx) uses W[ci, c, ry+p-s*dy, rx+p-s*dx]
    strided conv (Conv2d k, s, p)      = 3x3 conv  P*F -> F   tap (dy, dx), phase (ry, rx) uses W[co, ci, s*dy+ry+p, s*dx+rx+p]
    1x1 conv over concatenated HR maps = 1x1 conv over the image [n, h, w*P, F] (a 1x1 conv has no spatial structure)

(kernel taps outside [0, k) are zero entries of the packed operand), the two are each other's data gradients, and
torch.cat is free: concatenated maps are image-stacked slots of one tensor, read as several K sources of one launch.
PReLU (one slope per activation) is fused into the conv epilogue in inference; training stores the pre-activation and
runs the two stream kernels of csrc/drf_kernels.cu.  The frame recurrence (hidden_state, drf_net.py:42-45) is
sequential over T; _InBlock and _OutBlock run once over all T*n frames.  Backward = full BPTT over the T frames,
weight gradients batched over T*n images per conv.  torch carries device memory and streams; there is no CPU /
PyTorch fallback.
"""
import ctypes as C
import gc

import numpy as np
import torch

from . import lib as L
from . import ops
from .edsr_engine import _Layer, flatten_module_parameters, up_factors

PROJECTION = {2: (6, 2, 2), 3: (7, 3, 2), 4: (8, 4, 2), 8: (12, 8, 2)}     # drf_net.py:69-76
MAX_DY = 32                                                               # dY chunks per wgrad descriptor (<= PVSR_MAX_DY)


def _tile(n):
    """(bn, n_tiles_n) of an EPI_STORE launch with n output columns."""
    if n % 256 == 0:
        return 256, n // 256
    if n % 144 == 0:
        return 144, n // 144
    return 64, n // 64


class _Operand:
    """A packed weight operand, split into column groups of at most MAX_DY * 64 columns (the conv launch stages at most
    2304 biases, a wgrad descriptor takes at most PVSR_MAX_DY dY chunks): per group the gather index and the bf16 rows."""

    def __init__(self, table, dev):
        n_kb, n_total, _ = table.shape
        self.n_kb, self.n_total = n_kb, n_total
        chunks = n_total // 64
        self.groups = []
        for c0 in range(0, chunks, MAX_DY):
            c1 = min(c0 + MAX_DY, chunks)
            idx = torch.from_numpy(np.ascontiguousarray(table[:, 64 * c0:64 * c1, :].reshape(-1))).to(dev)
            w = torch.empty(idx.numel() // 64, 64, dtype=torch.bfloat16, device=dev)
            self.groups.append((64 * c0, 64 * c1, idx, w))


# ---------------------------------------------------------------------------------------------------- operand tables
# A table is the gather index of a packed operand bf16 [K blocks][n_total][64]: element -> flat index into the fp32
# parameter (or -1 = structural zero).  K block order = (source, tap, 64-channel block), taps in raster order.
def table_pointwise(c_out, n_src, cs):
    """1x1 conv over n_src concatenated sources of cs channels: parameter [c_out, n_src*cs, 1, 1]."""
    kb = cs // 64
    src, cb, col, c = np.meshgrid(np.arange(n_src), np.arange(kb), np.arange(c_out), np.arange(64), indexing='ij')
    idx = col * (n_src * cs) + src * cs + cb * 64 + c
    return idx.reshape(n_src * kb, c_out, 64).astype(np.int32)


def table_pointwise_T(c_out, n_src, cs, j):
    """Data gradient of the same conv wrt source j: K = c_out output channels, columns = the cs channels of source j."""
    kb = c_out // 64
    cb, col, c = np.meshgrid(np.arange(kb), np.arange(cs), np.arange(64), indexing='ij')
    idx = (cb * 64 + c) * (n_src * cs) + j * cs + col
    return idx.reshape(kb, cs, 64).astype(np.int32)


def table_expand(F, k, s, p):
    """F -> P*F conv: K channel = parameter dim 0, column (q, ch) = parameter dim 1, tap ky = ry + p - s*dy."""
    kb, P = F // 64, s * s
    tap, cb, q, ch, c = np.meshgrid(np.arange(9), np.arange(kb), np.arange(P), np.arange(F), np.arange(64), indexing='ij')
    ky = q // s + p - s * (tap // 3 - 1)
    kx = q % s + p - s * (tap % 3 - 1)
    idx = (((cb * 64 + c) * F + ch) * k + ky) * k + kx
    idx = np.where((ky >= 0) & (ky < k) & (kx >= 0) & (kx < k), idx, -1)
    return idx.reshape(9 * kb, P * F, 64).astype(np.int32)


def table_reduce(F, k, s, p):
    """P*F -> F conv: K channel (q, ch) = parameter dim 1, column = parameter dim 0, tap ky = s*dy + ry + p."""
    kb, P = F // 64, s * s
    tap, q, cb, col, c = np.meshgrid(np.arange(9), np.arange(P), np.arange(kb), np.arange(F), np.arange(64), indexing='ij')
    ky = s * (tap // 3 - 1) + q // s + p
    kx = s * (tap % 3 - 1) + q % s + p
    idx = ((col * F + cb * 64 + c) * k + ky) * k + kx
    idx = np.where((ky >= 0) & (ky < k) & (kx >= 0) & (kx < k), idx, -1)
    return idx.reshape(9 * P * kb, F, 64).astype(np.int32)


def projection_pairs(kind, k, s, p):
    """[(tap, phase q, ky, kx)] of the non-zero (tap, phase) blocks of a phase-stacked projection conv ('expand' =
    deconv, 'reduce' = strided conv): exactly k*k pairs, one per kernel tap of the k x k parameter."""
    out = []
    for tap in range(9):
        dy, dx = tap // 3 - 1, tap % 3 - 1
        for q in range(s * s):
            ry, rx = q // s, q % s
            ky, kx = (ry + p - s * dy, rx + p - s * dx) if kind == 'expand' else (s * dy + ry + p, s * dx + rx + p)
            if 0 <= ky < k and 0 <= kx < k:
                out.append((tap, q, ky, kx))
    return out


class _Node:
    """One Conv2d / ConvTranspose2d (+ its PReLU) of the LR-grid part of the net.

    kind: 'in1' (3x3, 1 -> 4F), 'pw' (1x1 over n_src sources of cs channels -> F), 'expand' (deconv), 'reduce' (strided)."""

    def __init__(self, eng, name, prelu, kind, n_src=1, cs=None):
        F, dev = eng.F, eng.device
        k, s, p = eng.proj
        P = s * s
        self.name, self.prelu, self.kind, self.n_src = name, prelu, kind, n_src
        self.taps, self.k16_last = 9, 4
        self.dgrad = []                                   # per source: (packed operand table, columns)
        if kind == 'in1':
            sp = ops._spec(4 * F, 1, 3, 1, [0], 1, 1, 9, 4 * F)
            fwd, bias = ops.pack_index(sp).reshape(9, 4 * F, 64), np.arange(4 * F)
            self.kb, self.n_total, self.k16_last = 1, 4 * F, 1
        elif kind == 'pw':
            fwd, bias = table_pointwise(F, n_src, cs), np.arange(F)
            self.kb, self.n_total, self.taps, self.cs = cs // 64, F, 1, cs
            self.dgrad = [table_pointwise_T(F, n_src, cs, j) for j in range(n_src)]
        elif kind == 'expand':
            fwd, bias = table_expand(F, k, s, p), np.tile(np.arange(F), P)
            self.kb, self.n_total = F // 64, P * F
            self.dgrad = [table_reduce(F, k, s, p)]
        else:
            fwd, bias = table_reduce(F, k, s, p), np.arange(F)
            self.kb, self.n_total = P * F // 64, F
            self.dgrad = [table_expand(F, k, s, p)]
        self.n_kb = fwd.shape[0]
        self.fwd = _Operand(fwd.reshape(self.n_kb, self.n_total, 64), dev)
        self.bwd = [_Operand(t, dev) for t in self.dgrad]
        self.idx_b = torch.from_numpy(bias.astype(np.int32)).to(dev)
        self.b = torch.empty(self.n_total, dtype=torch.float32, device=dev)


class _Geometry:
    """Buffers of one (T, n, h, w, train) input geometry."""

    def __init__(self, eng, T, n, h, w, train):
        self.T, self.n, self.h, self.w, self.train = T, n, h, w, train
        F, G, dev = eng.F, eng.G, eng.device
        P = eng.proj[1] ** 2
        Ts = T if train else 1                     # frames kept of the feedback block's intermediates
        self.Ts = Ts
        bf = dict(dtype=torch.bfloat16, device=dev)

        def stack(slots, ch):
            return torch.empty(slots * n, h, w, ch, **bf)

        self.x32 = torch.empty(T * n, h, w, dtype=torch.float32, device=dev)
        names = dict(x64=(T, 64), u=(T, 4 * F), a=(T, F), hid=(T + 1, F), lr=((G + 1) * Ts, F), m=(G * Ts, F),
                     hr=(G * Ts, P * F), d=(G * Ts, P * F), feat=(T, F))
        self.y = {k: stack(*v) for k, v in names.items()}
        if train:
            pre = dict(u=(T, 4 * F), a=(T, F), f=(T, F), lr=((G + 1) * T, F), m=(G * T, F), hr=(G * T, P * F),
                       d=(G * T, P * F))
            self.z = {k: stack(*v) for k, v in pre.items()}
            # gradients wrt the post-activation tensors (accumulated over consumers), one zero-able allocation
            gn = dict(u=(T, 4 * F), a=(T, F), hid=(T + 1, F), lr=((G + 1) * T, F), m=(G * T, F), hr=(G * T, P * F),
                      d=(G * T, P * F), feat=(T, F))
            total = sum(s * n * h * w * ch for s, ch in gn.values())
            self.g_all = torch.zeros(total, **bf)
            self.g, off = {}, 0
            for k, (s, ch) in gn.items():
                cnt = s * n * h * w * ch
                self.g[k] = self.g_all[off:off + cnt].view(s * n, h, w, ch)
                off += cnt
        self.sizes = [(h, w)]
        for r in eng.factors:
            self.sizes.append((self.sizes[-1][0] * r, self.sizes[-1][1] * r))
        self.up = [torch.empty(T * n, hh, ww, F, **bf) for hh, ww in self.sizes[1:]]
        H, W = self.sizes[-1]
        self.out16 = torch.empty(T * n, H, W, 16, dtype=torch.float32, device=dev)
        self.out = torch.empty(T, n, 1, H, W, dtype=torch.float32, device=dev)
        self.graph_fwd = self.graph_bwd = None
        if train:
            self.dout = torch.zeros(T, n, 1, H, W, dtype=torch.float32, device=dev)
            self.target = torch.empty(T, n, 1, H, W, dtype=torch.float32, device=dev)
            self.loss = torch.zeros((), dtype=torch.float32, device=dev)
            self.g64 = torch.empty(T * n, H, W, 64, **bf)
            self.dup = [torch.empty_like(u) for u in self.up]
            self.dw_off, total = {}, 0
            for nd in eng.nodes:
                self.dw_off[nd.name] = total
                total += nd.n_kb * nd.n_total * 64 + nd.n_total
            for l in eng.head_layers:
                self.dw_off[l.name] = total
                total += l.n_kb * l.n_total * 64 + l.n_total
            self.dw = torch.zeros(total, dtype=torch.float32, device=dev)
            self.wg_launches = None
            self.scatter_table = None
            self.jobs_ready = False


class DRFEngine:
    def __init__(self, net):
        self.net = net
        self.F, self.G, self.s = net.num_features, net.num_groups, net.upscale_factor
        self.proj = PROJECTION[self.s]
        self.factors = up_factors(self.s)
        self.use_graph = True
        self.device = None
        self.nodes = None
        self.geoms = {}
        self._idx = {}
        self._packed_version = None
        self._flat = None

    # ------------------------------------------------------------------------------------------ parameters
    def index(self, spec, bias):          # used by edsr_engine._Layer (the _OutBlock convs)
        key = (bytes(spec), bias)
        t = self._idx.get(key)
        if t is None:
            t = torch.from_numpy(ops.pack_bias_index(spec) if bias else ops.pack_index(spec)).to(self.device)
            self._idx[key] = t
        return t

    def _build(self, device):
        if self.nodes is not None and self.device == device:
            return
        self.device = device
        self._idx, self.geoms, self._packed_version = {}, {}, None
        F, G = self.F, self.G
        N = {}

        def add(key, *a, **kw):
            N[key] = _Node(self, *a, **kw)

        add('in1', 'in_block.conv1', 'in_block.prelu1', 'in1')
        add('in2', 'in_block.conv2', 'in_block.prelu2', 'pw', 1, 4 * F)
        add('fin', 'f_block.in_block.conv', 'f_block.in_block.prelu', 'pw', 2, F)
        for i in range(G):
            u, d = f'f_block.up_blocks.{i}.', f'f_block.down_blocks.{i}.'
            if i == 0:
                add(('up', 0), u + 'deconv', u + 'prelu', 'expand')
                add(('down', 0), d + 'conv', d + 'prelu', 'reduce')
            else:
                add(('upc', i), u + 'conv1', u + 'prelu1', 'pw', i + 1, F)
                add(('up', i), u + 'deconv2', u + 'prelu2', 'expand')
                add(('downc', i), d + 'conv1', d + 'prelu1', 'pw', i + 1, F)
                add(('down', i), d + 'conv2', d + 'prelu2', 'reduce')
        add('fout', 'f_block.out_block.conv', 'f_block.out_block.prelu', 'pw', G, F)
        self.N = N
        self.nodes = list(N.values())
        self.head_layers = [_Layer(self, f'out_block.conv{k + 1}', F, F * r * r, 'up', r)
                            for k, r in enumerate(self.factors)]
        self.head_layers.append(_Layer(self, f'out_block.conv{len(self.factors) + 1}', F, 1, 'tail'))

    def _named(self):
        return dict(self.net.named_parameters())

    def _param_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.net.parameters())

    def params_changed(self):
        self._packed_version = None

    def flatten_parameters(self):
        """One flat fp32 parameter buffer + one flat gradient buffer (single all-reduce / single Adam kernel), as
        RefineNetEngine.flatten_parameters."""
        if self._flat is None:
            self._flat = flatten_module_parameters(self.net)
            self.params_changed()
        return self._flat

    def _upload_table(self, jobs):
        arr = (L.TableJob * len(jobs))(*jobs)
        t = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        return t, max(j.n for j in jobs)

    def _ensure_packed(self):
        """fp32 master parameters -> bf16 forward / data-gradient operands + packed biases: ONE table-driven launch."""
        ver = self._param_version()
        if self._packed_version == ver:
            return
        P = self._named()
        key = tuple(p.data_ptr() for p in P.values())
        if getattr(self, '_pack_key', None) != key:
            jobs = []
            for nd in self.nodes:
                w, b = P[nd.name + '.weight'], P[nd.name + '.bias']
                if w.dtype != torch.float32 or not w.is_contiguous() or b.dtype != torch.float32:
                    raise L.PvsrError(f'{nd.name}: parameters must be contiguous fp32')
                jobs.append(L.TableJob(b.data_ptr(), nd.idx_b.data_ptr(), nd.b.data_ptr(), nd.idx_b.numel(), 1.0, L.TJ_GATHER))
                for op in [nd.fwd] + nd.bwd:
                    for _, _, idx, wp in op.groups:
                        jobs.append(L.TableJob(w.data_ptr(), idx.data_ptr(), wp.data_ptr(), idx.numel(), 1.0, L.TJ_PACK))
            for l in self.head_layers:
                w, b = P[l.name + '.weight'], P[l.name + '.bias']
                jobs.append(L.TableJob(w.data_ptr(), l.idx_w.data_ptr(), l.w.data_ptr(), l.idx_w.numel(), 1.0, L.TJ_PACK))
                jobs.append(L.TableJob(b.data_ptr(), l.idx_b.data_ptr(), l.b.data_ptr(), l.idx_b.numel(), 1.0, L.TJ_GATHER))
                jobs.append(L.TableJob(w.data_ptr(), l.idx_wt.data_ptr(), l.wt.data_ptr(), l.idx_wt.numel(), 1.0, L.TJ_PACK))
            self._pack_table, self._pack_max = self._upload_table(jobs)
            self._pack_jobs, self._pack_key = len(jobs), key
        L.check(L.load().pvsr_run_table(L.ptr(self._pack_table), self._pack_jobs, self._pack_max, L.current_stream()),
                'pack table')
        self._packed_version = ver

    # ------------------------------------------------------------------------------------------ geometry
    def geometry(self, inputs, train):
        x0 = inputs[0]
        if not x0.is_cuda:
            raise L.PvsrError('DRFNet (B200) runs on CUDA only; there is no CPU fallback - move inputs to cuda')
        if x0.dim() != 4 or x0.shape[1] != 1:
            raise ValueError(f'expected inputs of shape (N, 1, h, w), got {tuple(x0.shape)}')
        self._build(x0.device)
        n, _, h, w = x0.shape
        key = (len(inputs), n, h, w, bool(train))
        g = self.geoms.get(key)
        if g is None:
            g = _Geometry(self, len(inputs), n, h, w, train)
            self.geoms[key] = g
        return g

    # ------------------------------------------------------------------------------------------ launches
    def _phase(self, t):
        """[imgs, h, w, P*F] seen as the image [imgs, h, w*P, F] (1x1 convs over HR maps)."""
        return t.view(t.shape[0], t.shape[1], t.shape[2] * (t.shape[3] // self.F), self.F)

    def _conv(self, g, nd, views, srcs, n_img, y, z=None):
        """y = PReLU(conv + bias).  Inference: fused epilogue.  Training: z = conv + bias stored, y by the stream kernel."""
        slope = self._named()[nd.prelu + '.weight']
        dst = y if z is None else z
        for c0, c1, _, wp in nd.fwd.groups:
            bn, ntn = _tile(c1 - c0)
            ops.conv3x3(views, srcs, n_img, wp, bn, epi=L.EPI_STORE, kb_per_src=nd.kb, k16_last=nd.k16_last, taps=nd.taps,
                        bias=nd.b[c0:c1], n_tiles_n=ntn, out_hw=(y.shape[1], y.shape[2]), out_bf16=dst[..., c0:c1],
                        out_ch=dst.shape[3], prelu=slope if z is None else None)
        if z is not None:
            L.check(L.load().pvsr_prelu_fwd_bf16(L.ptr(z), L.ptr(slope), L.ptr(y), y.numel(), L.current_stream()),
                    'prelu_fwd')

    def _slot(self, g, t, name, slot, frame, frames):
        """Images [n] of slot `slot`, frame `frame` of a stack with `frames` frames per slot."""
        b = (slot * frames + frame) * g.n
        return t[name][b:b + g.n]

    def _forward_launches(self, g):
        lib, st = L.load(), L.current_stream()
        N, G, n, T, Ts, F = self.N, self.G, g.n, g.T, g.Ts, self.F
        Y, Z = g.y, (g.z if g.train else None)
        TN = T * n
        L.check(lib.pvsr_pad_channel_bf16(L.ptr(g.x32), L.ptr(Y['x64']), TN * g.h * g.w, st), 'pad_channel')
        # _InBlock over all frames (drf_net.py:52-58)
        self._conv(g, N['in1'], [(Y['x64'], 1)], [0], TN, Y['u'], Z['u'] if Z else None)
        self._conv(g, N['in2'], [(Y['u'], 1)], [0], TN, Y['a'], Z['a'] if Z else None)
        Y['hid'][:n].copy_(Y['a'][:n])                                 # hidden state of frame 0 = its own features (:42-43)
        for t in range(T):
            tf = t if g.train else 0

            def lr(j, src=Y):
                return self._slot(g, src, 'lr', j, tf, Ts)

            def zz(name, j):
                return self._slot(g, Z, name, j, tf, Ts) if Z else None

            # _FBlock.in_block on cat(input, hidden) (:119-120)
            self._conv(g, N['fin'], [(Y['a'], 1), (Y['hid'], 1)], [(0, t * n, 0, 0, 0), (1, t * n, 0, 0, 0)], n, lr(0),
                       zz('lr', 0))
            for i in range(G):                                          # projection groups (:123-129)
                if i == 0:
                    src_views, src_list = [(Y['lr'], 1)], [(0, tf * n, 0, 0, 0)]
                else:
                    m = self._slot(g, Y, 'm', i, tf, Ts)
                    self._conv(g, N[('upc', i)], [(Y['lr'], 1)], [(0, (j * Ts + tf) * n, 0, 0, 0) for j in range(i + 1)], n,
                               m, zz('m', i))
                    src_views, src_list = [(Y['m'], 1)], [(0, (i * Ts + tf) * n, 0, 0, 0)]
                hr = self._slot(g, Y, 'hr', i, tf, Ts)
                self._conv(g, N[('up', i)], src_views, src_list, n, hr, zz('hr', i))
                if i == 0:
                    src_views, src_list = [(Y['hr'], 1)], [(0, tf * n, 0, 0, 0)]
                else:
                    d = self._slot(g, Y, 'd', i, tf, Ts)
                    zd = zz('d', i)
                    self._conv(g, N[('downc', i)], [(self._phase(Y['hr']), 1)],
                               [(0, (j * Ts + tf) * n, 0, 0, 0) for j in range(i + 1)], n, self._phase(d),
                               self._phase(zd) if zd is not None else None)
                    src_views, src_list = [(Y['d'], 1)], [(0, (i * Ts + tf) * n, 0, 0, 0)]
                self._conv(g, N[('down', i)], src_views, src_list, n, lr(i + 1), zz('lr', i + 1))
            f = Y['hid'][(t + 1) * n:(t + 2) * n]
            self._conv(g, N['fout'], [(Y['lr'], 1)], [(0, (j * Ts + tf) * n, 0, 0, 0) for j in range(1, G + 1)], n, f,
                       Z['f'][t * n:(t + 1) * n] if Z else None)
            # global residual skip (:46)
            L.check(lib.pvsr_add_bf16(L.ptr(Y['a'][t * n:(t + 1) * n]), L.ptr(f), L.ptr(Y['feat'][t * n:(t + 1) * n]),
                                      f.numel(), st), 'add_bf16')
        # _OutBlock over all frames (:136-147)
        src = Y['feat']
        for k, r in enumerate(self.factors):
            l = self.head_layers[k]
            bn = 256 if r == 2 else 192
            ops.conv3x3(src, [0], TN, l.w, bn, epi=L.EPI_PS, kb_per_src=l.fwd.kb_per_src, bias=l.b,
                        n_tiles_n=l.n_total // bn, out_bf16=g.up[k], ps_r=r)
            src = g.up[k]
        l = self.head_layers[-1]
        ops.conv3x3(src, [0], TN, l.w, 64, epi=L.EPI_STORE, kb_per_src=l.fwd.kb_per_src, bias=l.b, n_tiles_n=1,
                    out_f32=g.out16, out_ch=16, n_store=16)
        H, W = g.sizes[-1]
        L.check(lib.pvsr_take_channel0_f32(L.ptr(g.out16), 16, L.ptr(g.out), TN * H * W, st), 'take_channel0')

    # ------------------------------------------------------------------------------------------ backward
    def _prelu_bwd(self, nd, gy, z, grads):
        L.check(L.load().pvsr_prelu_bwd_bf16(L.ptr(gy), L.ptr(z), L.ptr(self._named()[nd.prelu + '.weight']), L.ptr(gy),
                                             L.ptr(grads[nd.prelu + '.weight']), gy.numel(), L.current_stream()),
                'prelu_bwd')

    def _dgrad(self, nd, j, dz, n_img, dst, accumulate=True):
        """dst (+)= data gradient of node `nd` wrt its source j, from dz [n_img, h', w', n_total]."""
        kb = dz.shape[3] // 64
        for c0, c1, _, wp in nd.bwd[j].groups:
            bn, ntn = _tile(c1 - c0)
            part = dst[..., c0:c1]
            ops.conv3x3([(dz, 1)], [0], n_img, wp, bn, epi=L.EPI_STORE, kb_per_src=kb, taps=nd.taps, n_tiles_n=ntn,
                        out_bf16=part, res=part if accumulate else None, out_ch=dst.shape[3],
                        out_hw=(dst.shape[1], dst.shape[2]))

    def _backward_launches(self, g, grads):
        lib, st = L.load(), L.current_stream()
        N, G, n, T, F = self.N, self.G, g.n, g.T, self.F
        Y, Z, Gd = g.y, g.z, g.g
        TN = T * n
        H, W = g.sizes[-1]
        g.dw.zero_()
        g.g_all.zero_()
        # ---- _OutBlock, all frames
        L.check(lib.pvsr_pad_channel_bf16(L.ptr(g.dout), L.ptr(g.g64), TN * H * W, st), 'pad_channel(dout)')
        tail = self.head_layers[-1]
        bn, ntn = _tile(F)
        ops.conv3x3(g.g64, [0], TN, tail.wt, bn, epi=L.EPI_STORE, kb_per_src=1, k16_last=1, n_tiles_n=ntn,
                    out_bf16=g.dup[-1])
        for k in reversed(range(len(self.factors))):
            r = self.factors[k]
            l = self.head_layers[k]
            d_in = g.dup[k - 1] if k > 0 else Gd['feat']
            ops.conv3x3([(g.dup[k], r)], [(0, 0, 0, q % r, q // r) for q in range(r * r)], TN, l.wt, bn,
                        epi=L.EPI_STORE, kb_per_src=l.bwd.kb_per_src, n_tiles_n=ntn, out_bf16=d_in,
                        out_hw=(d_in.shape[1], d_in.shape[2]))
        # features = in_features + f_features (:46): the gradient reaches both
        Gd['a'].copy_(Gd['feat'])
        Gd['hid'][n:].copy_(Gd['feat'])
        # ---- feedback block, frames in reverse (BPTT through hidden_state)
        for t in reversed(range(T)):
            def sl(stack, name, j):
                return self._slot(g, stack, name, j, t, T)

            gf = Gd['hid'][(t + 1) * n:(t + 2) * n]
            self._prelu_bwd(N['fout'], gf, Z['f'][t * n:(t + 1) * n], grads)
            for j in range(G):
                self._dgrad(N['fout'], j, gf, n, sl(Gd, 'lr', j + 1))
            for i in reversed(range(G)):
                glr = sl(Gd, 'lr', i + 1)
                self._prelu_bwd(N[('down', i)], glr, sl(Z, 'lr', i + 1), grads)
                if i == 0:
                    self._dgrad(N[('down', 0)], 0, glr, n, sl(Gd, 'hr', 0))
                else:
                    gd = sl(Gd, 'd', i)
                    self._dgrad(N[('down', i)], 0, glr, n, gd)
                    self._prelu_bwd(N[('downc', i)], gd, sl(Z, 'd', i), grads)
                    for j in range(i + 1):
                        self._dgrad(N[('downc', i)], j, self._phase(gd), n, self._phase(sl(Gd, 'hr', j)))
                ghr = sl(Gd, 'hr', i)
                self._prelu_bwd(N[('up', i)], ghr, sl(Z, 'hr', i), grads)
                if i == 0:
                    self._dgrad(N[('up', 0)], 0, ghr, n, sl(Gd, 'lr', 0))
                else:
                    gm = sl(Gd, 'm', i)
                    self._dgrad(N[('up', i)], 0, ghr, n, gm)
                    self._prelu_bwd(N[('upc', i)], gm, sl(Z, 'm', i), grads)
                    for j in range(i + 1):
                        self._dgrad(N[('upc', i)], j, gm, n, sl(Gd, 'lr', j))
            gl0 = sl(Gd, 'lr', 0)
            self._prelu_bwd(N['fin'], gl0, sl(Z, 'lr', 0), grads)
            self._dgrad(N['fin'], 0, gl0, n, Gd['a'][t * n:(t + 1) * n])
            self._dgrad(N['fin'], 1, gl0, n, Gd['hid'][t * n:(t + 1) * n])
        # hidden state of frame 0 was in_features of frame 0
        L.check(lib.pvsr_add_bf16(L.ptr(Gd['a'][:n]), L.ptr(Gd['hid'][:n]), L.ptr(Gd['a'][:n]), Gd['a'][:n].numel(), st),
                'add_bf16')
        # ---- _InBlock, all frames
        self._prelu_bwd(N['in2'], Gd['a'], Z['a'], grads)
        self._dgrad(N['in2'], 0, Gd['a'], TN, Gd['u'], accumulate=False)
        self._prelu_bwd(N['in1'], Gd['u'], Z['u'], grads)
        # ---- weight gradients (batched over all T*n images) and the scatter into the parameter gradients
        for arr, n_desc, _ in g.wg_launches:
            L.check(lib.pvsr_conv3x3_wgrad_multi(C.cast(arr, C.c_void_p), n_desc, 0, st), 'wgrad')
        L.check(lib.pvsr_run_table(L.ptr(g.scatter_table), g.scatter_jobs, g.scatter_max, st), 'scatter table')

    def _wg_desc(self, g, views, srcs, dys, out_hw, n_img, kb, taps, n_total, dw_elem_off, db_elem_off,
                 with_bias=1):
        d = L.WgradDesc()
        d.H, d.W = out_hw
        d.n_img = n_img
        d.n_views = len(views)
        for i, (t, mul) in enumerate(views):
            d.views[i].ptr = t.data_ptr()
            d.views[i].channels = t.shape[3]
            d.views[i].H, d.views[i].W = t.shape[1], t.shape[2]
            d.views[i].images = t.shape[0]
            d.views[i].mul = mul
        d.n_src = len(srcs)
        for i, (v, base, ch0, ox, oy) in enumerate(srcs):
            d.src_view[i], d.src_img_base[i], d.src_ch0[i], d.src_off_x[i], d.src_off_y[i] = v, base, ch0, ox, oy
        d.n_dy = len(dys)
        for i, (v, base, ch0, ox, oy) in enumerate(dys):
            d.dy_view[i], d.dy_img_base[i], d.dy_ch0[i], d.dy_off_x[i], d.dy_off_y[i] = v, base, ch0, ox, oy
        d.kb_per_src, d.taps, d.n_total, d.with_bias, d.n_splits = kb, taps, n_total, with_bias, 0
        d.dw_packed = g.dw.data_ptr() + 4 * dw_elem_off
        d.db_packed = g.dw.data_ptr() + 4 * db_elem_off
        return d

    def _build_wgrad_launches(self, g, grads):
        """One wgrad launch per conv over all T*n images (X = the forward sources, dY = the PReLU-adjointed gradient of
        the conv's output) + the scatter table packed gradient -> parameter gradient."""
        lib, st = L.load(), L.current_stream()
        N, G, n, T, F = self.N, self.G, g.n, g.T, self.F
        Y, Gd = g.y, g.g
        TN = T * n
        sb = lib.pvsr_wgrad_scratch_bytes()
        launches, scatter = [], []
        base = g.dw.data_ptr()

        def add(descs):
            arr = (L.WgradDesc * len(descs))(*descs)
            scratch = torch.empty(sb, dtype=torch.uint8, device=self.device)
            arr[0].job_scratch = scratch.data_ptr()
            L.check(lib.pvsr_conv3x3_wgrad_multi(C.cast(arr, C.c_void_p), len(descs), 1, st), 'wgrad job upload')
            launches.append((arr, len(descs), scratch))

        def node_wgrad(nd, x_view, srcs, dy_t, out_hw):
            """x_view: X tensor; srcs: [(img_base)] per source; dy_t: gradient tensor whose slot layout matches."""
            off = g.dw_off[nd.name]
            n_w = nd.n_kb * nd.n_total * 64
            descs, o = [], off
            for c0, c1, idx, _ in nd.fwd.groups:
                cols = c1 - c0
                dys = [(1, dy_t[1], c, 0, 0) for c in range(c0, c1, 64)]
                descs.append(self._wg_desc(g, [(x_view, 1), (dy_t[0], 1)],
                                           [(0, b, 0, 0, 0) for b in srcs], dys, out_hw, TN, nd.kb, nd.taps, cols,
                                           o, off + n_w + c0))
                scatter.append(L.TableJob(base + 4 * o, idx.data_ptr(), grads[nd.name + '.weight'].data_ptr(),
                                          nd.n_kb * cols * 64, 1.0, L.TJ_SCATTER))
                o += nd.n_kb * cols * 64
            for d in descs:                      # separate launches: descriptors of one launch share one dw base only
                add([d])
            scatter.append(L.TableJob(base + 4 * (off + n_w), nd.idx_b.data_ptr(), grads[nd.name + '.bias'].data_ptr(),
                                      nd.n_total, 1.0, L.TJ_SCATTER))

        def projection_wgrad(nd, x_t, x_base, dz_t, dz_base):
            """Weight gradient of a phase-stacked projection conv WITHOUT its structural zeros: one (tap, phase) block per
            kernel tap of the k x k parameter (k*k of the 9 s*s blocks the dense 3x3 form reduces), as single-tap sources
            shifted by the tap offset.  All descriptors of a node ride in one launch."""
            k, s_, p_ = self.proj
            kbF = F // 64
            off = g.dw_off[nd.name]
            n_w = nd.n_kb * nd.n_total * 64
            pairs = projection_pairs(nd.kind, k, s_, p_)
            views = [(x_t, 1), (dz_t, 1)]
            descs, o, ob = [], off, off + n_w
            wname, bname = grads[nd.name + '.weight'].data_ptr(), grads[nd.name + '.bias'].data_ptr()
            cb, c64 = np.meshgrid(np.arange(kbF), np.arange(64), indexing='ij')
            kch = (cb * 64 + c64)                                    # K channel of (cb, c)
            if nd.kind == 'reduce':
                # X = phase-stacked map (source = phase q's F channels shifted by the tap), dY = dz (F columns)
                for g0 in range(0, len(pairs), 10):
                    grp = pairs[g0:g0 + 10]
                    srcs = [(0, x_base, q * F, tap % 3 - 1, tap // 3 - 1) for tap, q, _, _ in grp]
                    dys = [(1, dz_base, 64 * c, 0, 0) for c in range(kbF)]
                    first = g0 == 0
                    descs.append(self._wg_desc(g, views, srcs, dys, hw, TN, kbF, 1, F, o, ob, with_bias=int(first)))
                    idx = np.empty((len(grp), kbF, F, 64), dtype=np.int64)
                    col = np.arange(F)[None, :, None]
                    for si, (_, _, ky, kx) in enumerate(grp):
                        idx[si] = ((col * F + kch[:, None, :]) * k + ky) * k + kx
                    t = torch.from_numpy(idx.reshape(-1).astype(np.int32)).to(self.device)
                    keep.append(t)
                    scatter.append(L.TableJob(base + 4 * o, t.data_ptr(), wname, t.numel(), 1.0, L.TJ_SCATTER))
                    o += t.numel()
                scatter.append(L.TableJob(base + 4 * ob, nd.idx_b.data_ptr(), bname, F, 1.0, L.TJ_SCATTER))
            else:
                # X = the F-channel input shifted by the tap, dY = the phases of dz that tap reaches
                per = max(1, MAX_DY // kbF)
                for tap in range(9):
                    qs = [(q, ky, kx) for t_, q, ky, kx in pairs if t_ == tap]
                    for g0 in range(0, len(qs), per):
                        grp = qs[g0:g0 + per]
                        cols = len(grp) * F
                        srcs = [(0, x_base, 0, tap % 3 - 1, tap // 3 - 1)]
                        dys = [(1, dz_base, q * F + 64 * c, 0, 0) for q, _, _ in grp for c in range(kbF)]
                        centre = tap == 4                            # the centre tap reaches every phase once: bias gradient
                        descs.append(self._wg_desc(g, views, srcs, dys, hw, TN, kbF, 1, cols, o, ob,
                                                   with_bias=int(centre)))
                        idx = np.empty((kbF, len(grp), F, 64), dtype=np.int64)
                        ch = np.arange(F)[None, :, None]
                        for gi, (_, ky, kx) in enumerate(grp):
                            idx[:, gi] = ((kch[:, None, :] * F + ch) * k + ky) * k + kx
                        t = torch.from_numpy(idx.reshape(-1).astype(np.int32)).to(self.device)
                        keep.append(t)
                        scatter.append(L.TableJob(base + 4 * o, t.data_ptr(), wname, t.numel(), 1.0, L.TJ_SCATTER))
                        o += t.numel()
                        if centre:
                            tb = torch.from_numpy(np.tile(np.arange(F), len(grp)).astype(np.int32)).to(self.device)
                            keep.append(tb)
                            scatter.append(L.TableJob(base + 4 * ob, tb.data_ptr(), bname, cols, 1.0, L.TJ_SCATTER))
                            ob += cols
            assert o <= off + n_w and ob <= off + n_w + nd.n_total
            add(descs)

        keep = []
        hw = (g.h, g.w)
        P = self.proj[1] ** 2
        hwp = (g.h, g.w * P)
        node_wgrad(N['in1'], Y['x64'], [0], (Gd['u'], 0), hw)
        node_wgrad(N['in2'], Y['u'], [0], (Gd['a'], 0), hw)
        # fin: sources a[t], hid[t] for all t; dY = gradient of lr slot 0
        off = g.dw_off[N['fin'].name]
        nd = N['fin']
        d = self._wg_desc(g, [(Y['a'], 1), (Y['hid'], 1), (Gd['lr'], 1)],
                          [(0, 0, 0, 0, 0), (1, 0, 0, 0, 0)], [(2, 0, 64 * c, 0, 0) for c in range(F // 64)], hw, TN, nd.kb, 1,
                          F, off, off + nd.n_kb * F * 64)
        add([d])
        scatter.append(L.TableJob(base + 4 * off, nd.fwd.groups[0][2].data_ptr(), grads[nd.name + '.weight'].data_ptr(),
                                  nd.n_kb * F * 64, 1.0, L.TJ_SCATTER))
        scatter.append(L.TableJob(base + 4 * (off + nd.n_kb * F * 64), nd.idx_b.data_ptr(),
                                  grads[nd.name + '.bias'].data_ptr(), F, 1.0, L.TJ_SCATTER))
        for i in range(G):
            if i == 0:
                projection_wgrad(N[('up', 0)], Y['lr'], 0, Gd['hr'], 0)
                projection_wgrad(N[('down', 0)], Y['hr'], 0, Gd['lr'], TN)
            else:
                node_wgrad(N[('upc', i)], Y['lr'], [j * TN for j in range(i + 1)], (Gd['m'], i * TN), hw)
                projection_wgrad(N[('up', i)], Y['m'], i * TN, Gd['hr'], i * TN)
                node_wgrad(N[('downc', i)], self._phase(Y['hr']), [j * TN for j in range(i + 1)],
                           (self._phase(Gd['d']), i * TN), hwp)
                projection_wgrad(N[('down', i)], Y['d'], i * TN, Gd['lr'], (i + 1) * TN)
        node_wgrad(N['fout'], Y['lr'], [j * TN for j in range(1, G + 1)], (Gd['hid'], n), hw)
        # _OutBlock convs (as edsr_engine: tail against g64, up-samplers against pixel-unshuffled views of dup)
        H, W = g.sizes[-1]
        cb = F // 64
        tail = self.head_layers[-1]
        off = g.dw_off[tail.name]
        n_w = tail.n_kb * tail.n_total * 64
        add([self._wg_desc(g, [(g.up[-1], 1), (g.g64, 1)], [(0, 0, 0, 0, 0)], [(1, 0, 0, 0, 0)],
                           (H, W), TN, tail.fwd.kb_per_src, 9, tail.n_total, off, off + n_w)])
        for k, r in enumerate(self.factors):
            l = self.head_layers[k]
            x_in = g.up[k - 1] if k > 0 else Y['feat']
            off = g.dw_off[l.name]
            n_w = l.n_kb * l.n_total * 64
            dys = [(1, 0, c * 64, q % r, q // r) for q in range(r * r) for c in range(cb)]
            add([self._wg_desc(g, [(x_in, 1), (g.dup[k], r)], [(0, 0, 0, 0, 0)], dys, g.sizes[k], TN,
                               l.fwd.kb_per_src, 9, l.n_total, off, off + n_w)])
        for l in self.head_layers:
            off = g.dw_off[l.name]
            n_w = l.n_kb * l.n_total * 64
            scatter.append(L.TableJob(base + 4 * off, l.idx_w.data_ptr(), grads[l.name + '.weight'].data_ptr(), n_w, 1.0,
                                      L.TJ_SCATTER))
            scatter.append(L.TableJob(base + 4 * (off + n_w), l.idx_b.data_ptr(), grads[l.name + '.bias'].data_ptr(),
                                      l.n_total, 1.0, L.TJ_SCATTER))
        g.wg_launches = launches
        g.wg_keep = keep                        # scatter indices of the projection convs (referenced by the table)
        g.scatter_table, g.scatter_max = self._upload_table(scatter)
        g.scatter_jobs = len(scatter)

    # ------------------------------------------------------------------------------------------ public
    def _replay(self, g, which, fn):
        attr = 'graph_' + which
        state = getattr(g, attr)
        if not self.use_graph:
            fn()
            return
        if state is None:
            fn()
            setattr(g, attr, 1)
        elif state == 1:
            fn()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            # No garbage collection while capturing: a cycle collection that happens to run inside the capture may
            # destroy CUDA graphs / tensors of dead engines (cudaGraphExecDestroy, cudaFree), which CUDA forbids on a
            # capturing thread and which invalidates the capture (seen as a rare "operation not permitted when stream is
            # capturing").  thread_local: other threads (a DataLoader's pin-memory thread) may keep calling the runtime.
            gc_was_enabled = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                    fn()
            finally:
                if gc_was_enabled:
                    gc.enable()
            setattr(g, attr, graph)
        else:
            state.replay()

    def forward(self, inputs, train=False, clone=True):
        g = self.geometry(inputs, train)
        self._ensure_packed()
        for t, x in enumerate(inputs):
            g.x32[t * g.n:(t + 1) * g.n].copy_(x.reshape(g.n, g.h, g.w))
        if getattr(g, 'fwd_key', None) != self._slope_key():      # captured graphs hold the slope pointers
            g.fwd_key, g.graph_fwd = self._slope_key(), None
        self._replay(g, 'fwd', lambda: self._forward_launches(g))
        out = g.out.clone() if clone else g.out
        return [out[t] for t in range(g.T)], g

    def _slope_key(self):
        return tuple(p.data_ptr() for k, p in self.net.named_parameters() if 'prelu' in k)

    def backward(self, g, grads):
        """Backward of the last forward(train=True) of geometry `g`; g.dout holds dL/d(out)."""
        key = tuple(grads[k].data_ptr() for k in sorted(grads)) + self._slope_key()
        if getattr(g, 'bwd_key', None) != key:          # a captured graph writes to the buffers it was captured with
            self._build_wgrad_launches(g, grads)
            g.bwd_key, g.graph_bwd = key, None
        self._replay(g, 'bwd', lambda: self._backward_launches(g, grads))

    def grad_buffers(self):
        if getattr(self, '_grad_buf', None) is None:
            self._grad_buf = {k: torch.zeros_like(p) for k, p in self._named().items()}
        return self._grad_buf

    def loss_and_grads(self, inputs, targets, zero_grads=True):
        """Fused training-step body: forward, the frame-averaged nn.L1Loss of the VSR trainer
        (acdc_vsr_trainer.py:40-43,83-94) and backward into `p.grad`.  Returns (loss, outputs)."""
        outs, g = self.forward(inputs, train=True, clone=False)
        for t, tg in enumerate(targets):
            g.target[t].copy_(tg.reshape(g.target[t].shape))
        n = g.out.numel()
        w = getattr(self, '_lw', {}).get(n)
        if w is None:
            self._lw = getattr(self, '_lw', {})
            w = self._lw[n] = torch.tensor([1.0 / n], dtype=torch.float32, device=self.device)
        g.loss.zero_()
        L.check(L.load().pvsr_l1_multistage(L.ptr(g.out), L.ptr(g.target), L.ptr(w), 1, n, L.ptr(g.loss),
                                            L.ptr(g.dout), L.current_stream()), 'pvsr_l1_multistage')
        grads = {}
        for k, p in self._named().items():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads[k] = p.grad
        if zero_grads:
            if self._flat is not None:
                self._flat[1].zero_()
            else:
                for t in grads.values():
                    t.zero_()
        self.backward(g, grads)
        return g.loss.clone(), outs


class _DRFFunction(torch.autograd.Function):
    """Autograd bridge: `net(inputs)` in training mode is differentiable wrt the parameters, so the reference's VSR
    trainer sequence (acdc_vsr_trainer.py:40-47: any torch loss, loss.backward(), any torch optimiser) works unchanged;
    the input frames receive no gradient (the reference never asks for one)."""

    @staticmethod
    def forward(ctx, engine, T, names, *rest):
        inputs, params = rest[:T], rest[T:]
        outs, g = engine.forward(list(inputs), train=True, clone=True)
        g.fwd_serial = getattr(g, 'fwd_serial', 0) + 1
        ctx.engine, ctx.g, ctx.names, ctx.T, ctx.serial = engine, g, names, T, g.fwd_serial
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        engine, g = ctx.engine, ctx.g
        if g.fwd_serial != ctx.serial:
            raise L.PvsrError('the activations saved by this forward were overwritten by a later forward of the same shape (the plan keeps ONE set of training buffers per shape): call backward before the next forward')
        for t, go in enumerate(grad_outs):
            if go is None:
                g.dout[t].zero_()
            else:
                g.dout[t].copy_(go.reshape(g.dout[t].shape))
        bufs = engine.grad_buffers()
        for b in bufs.values():
            b.zero_()
        engine.backward(g, bufs)
        return (None, None, None) + (None,) * ctx.T + tuple(bufs[k].clone() for k in ctx.names)


def drf_train_forward(net, inputs):
    named = list(net.named_parameters())
    outs = _DRFFunction.apply(net.engine, len(inputs), tuple(k for k, _ in named), *[x.detach() for x in inputs],
                              *[p for _, p in named])
    return list(outs)
