"""Host-side runtime of EDSRNet (reference src/model/nets/edsr_net.py:8-71) on the RefineNet
conv core: every convolution of the net - head 1 -> F, 2R + 1 body convs F -> F, the up-sampler convs F -> r*r*F with
PixelShuffle as epilogue addressing, tail F -> 1 - and every data / weight gradient is a launch of the tcgen05
implicit-GEMM kernels behind `pvsr_conv3x3_fwd` / `pvsr_conv3x3_wgrad_staged` (include/pvsr.h):

    head        x (fp32, 1 ch)  -> channel 0 of a zero-padded 64-channel bf16 K block (pvsr_pad_channel_bf16), K = 16
    resblock i  t_i = relu(conv1(x_i))                      EPI_STORE, relu = 1
                x_{i+1} = res_scale * conv2(t_i) + x_i      EPI_STORE, out_scale = res_scale, res = x_i
    body.conv   u_0 = conv(x_R) + head                      EPI_STORE, res = head
    up-sampler  u_{k+1} = PixelShuffle_r(conv(u_k))         EPI_PS (columns grouped per sub-pixel)
    tail        out = conv(u_last)                          EPI_STORE, N = 64 tile of which 16 fp32 columns are stored,
                                                            column 0 is the image (pvsr_take_channel0_f32)
Backward mirrors it with transposed / tap-flipped packs of the same parameters (pack spec `transpose_flip`), the
ReLU adjoint as a mask on the stored t_i (`mask`), the residual adjoint as `res`, PixelShuffle's adjoint as
pixel-unshuffled TMA views, and one pixel-reduction wgrad launch per conv whose packed result is scattered into the
fp32 parameter-layout gradient (`pvsr_scatter_add[_scaled]`).  torch carries device memory, streams and (optionally)
replays the launch sequence as a CUDA graph; there is no CPU / PyTorch fallback.
"""
import ctypes as C
import gc
import math

import torch

from . import lib as L
from . import ops


def flatten_module_parameters(net):
    """Re-homes every parameter of `net` (and its .grad) as a view of ONE flat fp32 buffer each, 16-byte aligned per
    parameter; returns (flat_param, flat_grad).  Shared by the EDSR and DRFNet engines."""
    params = list(net.parameters())
    dev = params[0].device
    offs, total = [], 0
    for p in params:
        offs.append(total)
        total += (p.numel() + 3) // 4 * 4
    flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
    flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
    with torch.no_grad():
        for p, o in zip(params, offs):
            n = p.numel()
            flat_p[o:o + n].copy_(p.detach().reshape(-1))
            p.data = flat_p[o:o + n].view(p.shape)
            p.grad = flat_g[o:o + n].view(p.shape)
    return flat_p, flat_g


def up_factors(upscale_factor):
    if (math.log(upscale_factor, 2) % 1) == 0:
        return [2] * int(math.log(upscale_factor, 2))
    if upscale_factor == 3:
        return [3]
    raise NotImplementedError(f'upscale factor {upscale_factor}')


def _spec(c_out, c_in, n_src, src_ch, kb, n_total, ps_r=0, ps_ch=0, transpose_flip=0, k_ps_r=0):
    s = L.PackSpec()
    s.c_out, s.c_in, s.kh, s.kw, s.n_src = c_out, c_in, 3, 3, n_src
    s.src_ch, s.kb_per_src, s.taps, s.n_total = src_ch, kb, 9, n_total
    s.ps_r, s.ps_ch, s.transpose_flip, s.k_ps_r = ps_r, ps_ch, transpose_flip, k_ps_r
    return s


def _tile(n):
    """(bn, n_tiles_n) of an EPI_STORE launch with n output columns."""
    return (256, n // 256) if n % 256 == 0 else (64, n // 64)


class _Layer:
    """One Conv2d of the net: parameter names, pack specs / device indices, packed operands."""

    def __init__(self, eng, name, c_in, c_out, kind, r=0):
        self.name, self.c_in, self.c_out, self.kind, self.r = name, c_in, c_out, kind, r
        F = eng.F
        kb = max(1, c_in // 64)
        if kind == 'head':          # 1 -> F
            self.fwd = _spec(F, 1, 1, 1, 1, F)
            self.bwd = None         # no data gradient wrt the input image
        elif kind == 'body':        # F -> F
            self.fwd = _spec(F, F, 1, F, kb, F)
            self.bwd = _spec(F, F, 1, F, kb, F, transpose_flip=1)
        elif kind == 'up':          # F -> r*r*F, columns in pixel-shuffle order
            self.fwd = _spec(F * r * r, F, 1, F, kb, F * r * r, ps_r=r, ps_ch=F)
            self.bwd = _spec(F * r * r, F, r * r, F, kb, F, transpose_flip=1, k_ps_r=r)
        else:                       # tail F -> 1 (one real column of a 64-column tile)
            self.fwd = _spec(1, F, 1, F, kb, 64)
            self.bwd = _spec(1, F, 1, 1, 1, F, transpose_flip=1)
        self.n_kb = self.fwd.n_src * 9 * self.fwd.kb_per_src
        self.n_total = self.fwd.n_total
        self.idx_w = eng.index(self.fwd, False)
        self.idx_b = eng.index(self.fwd, True)
        self.idx_wt = eng.index(self.bwd, False) if self.bwd is not None else None
        dev = eng.device
        self.w = torch.empty(self.idx_w.numel() // 64, 64, dtype=torch.bfloat16, device=dev)
        self.b = torch.empty(self.n_total, dtype=torch.float32, device=dev)
        self.wt = (torch.empty(self.idx_wt.numel() // 64, 64, dtype=torch.bfloat16, device=dev)
                   if self.idx_wt is not None else None)


class _Geometry:
    """Buffers and wgrad job scratch of one (N, h, w, train) input geometry."""

    def __init__(self, eng, n, h, w, train):
        self.n, self.h, self.w, self.train = n, h, w, train
        F, dev = eng.F, eng.device
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.x32 = torch.empty(n, h, w, dtype=torch.float32, device=dev)
        self.x64 = torch.empty(n, h, w, 64, **bf)
        n_feat = (2 * eng.R + 2) if train else 3            # head, (t_i, x_{i+1}) per block, u_0 | ping-pong
        # ONE image-stacked tensor (slot-major, like the frame-major stacks of the RefineNet plan): in training, body
        # layer l reads slot l and writes slot l + 1, so a single TMA descriptor serves the weight gradients of all
        # body layers in one launch
        self.acts = torch.empty(n_feat * n, h, w, F, **bf)
        self.feat = [self.acts[i * n:(i + 1) * n] for i in range(n_feat)]
        self.sizes = [(h, w)]
        for r in eng.factors:
            self.sizes.append((self.sizes[-1][0] * r, self.sizes[-1][1] * r))
        self.up = [torch.empty(n, hh, ww, F, **bf) for hh, ww in self.sizes[1:]]
        H, W = self.sizes[-1]
        self.out16 = torch.empty(n, H, W, 16, dtype=torch.float32, device=dev)
        self.out = torch.empty(n, 1, H, W, dtype=torch.float32, device=dev)
        self.graph_fwd = self.graph_bwd = None
        if train:
            self.dout = torch.zeros(n, 1, H, W, dtype=torch.float32, device=dev)
            self.target = torch.empty(n, 1, H, W, dtype=torch.float32, device=dev)
            self.loss = torch.zeros((), dtype=torch.float32, device=dev)
            self.g64 = torch.empty(n, H, W, 64, **bf)
            self.dup = [torch.empty_like(u) for u in self.up]          # gradients wrt u_1 .. u_last
            # gradient stack, same slots as `acts`: slot j = dL/d(what acts slot j holds; pre-ReLU for the t_i slots)
            self.gacts = torch.empty(n_feat * n, h, w, F, **bf)
            self.gfeat = [self.gacts[i * n:(i + 1) * n] for i in range(n_feat)]
            # packed fp32 weight / bias gradients of every layer, side by side (zeroed once per step, scattered once)
            self.dw_off, total = {}, 0
            for l in eng.layers:
                self.dw_off[l.name] = total
                total += l.n_kb * l.n_total * 64 + l.n_total
            self.dw = torch.zeros(total, dtype=torch.float32, device=dev)
            self.wg_launches = None          # [(ctypes desc array, n_desc, job scratch tensor)]
            self.scatter_table = None
            self.jobs_ready = False


class EDSREngine:
    def __init__(self, net):
        self.net = net
        self.F, self.R = net.num_features, net.num_resblocks
        self.factors = up_factors(net.upscale_factor)
        self.res_scale = float(net.res_scale)
        self.use_graph = True
        self.device = None
        self.layers = None
        self.geoms = {}
        self._idx = {}
        self._packed_version = None
        self._flat = None

    # ------------------------------------------------------------------------------------------ parameters
    def index(self, spec, bias):
        key = (bytes(spec), bias)
        t = self._idx.get(key)
        if t is None:
            t = torch.from_numpy(ops.pack_bias_index(spec) if bias else ops.pack_index(spec)).to(self.device)
            self._idx[key] = t
        return t

    def _build_layers(self, device):
        if self.layers is not None and self.device == device:
            return
        self.device = device
        self._idx, self.geoms, self._packed_version = {}, {}, None
        F = self.F
        self.layers = [_Layer(self, 'head.0', 1, F, 'head')]
        for i in range(self.R):
            self.layers.append(_Layer(self, f'body.{i}.body.conv1', F, F, 'body'))
            self.layers.append(_Layer(self, f'body.{i}.body.conv2', F, F, 'body'))
        self.layers.append(_Layer(self, 'body.conv', F, F, 'body'))
        for k, r in enumerate(self.factors):
            self.layers.append(_Layer(self, f'tail.0.conv{k + 1}', F, F * r * r, 'up', r))
        self.layers.append(_Layer(self, 'tail.conv', F, 1, 'tail'))
        self.by_name = {l.name: l for l in self.layers}

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
        """fp32 master parameters -> bf16 forward / transposed operands + packed biases of every layer: ONE
        table-driven launch (pvsr_run_table) per optimiser step."""
        ver = self._param_version()
        if self._packed_version == ver:
            return
        P = self._named()
        key = tuple(p.data_ptr() for p in P.values())
        if getattr(self, '_pack_key', None) != key:
            jobs = []
            for l in self.layers:
                w, b = P[l.name + '.weight'], P[l.name + '.bias']
                if w.dtype != torch.float32 or not w.is_contiguous() or b.dtype != torch.float32:
                    raise L.PvsrError(f'{l.name}: parameters must be contiguous fp32')
                jobs.append(L.TableJob(w.data_ptr(), l.idx_w.data_ptr(), l.w.data_ptr(), l.idx_w.numel(), 1.0, L.TJ_PACK))
                jobs.append(L.TableJob(b.data_ptr(), l.idx_b.data_ptr(), l.b.data_ptr(), l.idx_b.numel(), 1.0,
                                       L.TJ_GATHER))
                if l.wt is not None:
                    jobs.append(L.TableJob(w.data_ptr(), l.idx_wt.data_ptr(), l.wt.data_ptr(), l.idx_wt.numel(), 1.0,
                                           L.TJ_PACK))
            self._pack_table, self._pack_max = self._upload_table(jobs)
            self._pack_jobs, self._pack_key = len(jobs), key
        L.check(L.load().pvsr_run_table(L.ptr(self._pack_table), self._pack_jobs, self._pack_max, L.current_stream()),
                'pack table')
        self._packed_version = ver

    # ------------------------------------------------------------------------------------------ geometry
    def geometry(self, x, train):
        if not x.is_cuda:
            raise L.PvsrError('EDSRNet (B200) runs on CUDA only; there is no CPU fallback - move inputs to cuda')
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f'expected an input of shape (N, 1, h, w), got {tuple(x.shape)}')
        self._build_layers(x.device)
        n, _, h, w = x.shape
        key = (n, h, w, bool(train))
        g = self.geoms.get(key)
        if g is None:
            g = _Geometry(self, n, h, w, train)
            self.geoms[key] = g
        return g

    # ------------------------------------------------------------------------------------------ launches
    def _conv(self, src, layer, out, relu=0, res=None, scale=0.0):
        bn, nt = _tile(self.F)
        ops.conv3x3(src, [0], src.shape[0], layer.w, bn, epi=L.EPI_STORE, kb_per_src=layer.fwd.kb_per_src,
                    k16_last=1 if layer.kind == 'head' else 4, bias=layer.b, n_tiles_n=nt, out_bf16=out, res=res,
                    relu=relu, out_scale=scale)

    def _forward_launches(self, g):
        lib, st = L.load(), L.current_stream()
        F, R, by = self.F, self.R, self.by_name
        n_px = g.n * g.h * g.w
        L.check(lib.pvsr_pad_channel_bf16(L.ptr(g.x32), L.ptr(g.x64), n_px, st), 'pad_channel')
        feat = g.feat
        self._conv(g.x64, by['head.0'], feat[0])
        # training keeps every activation (x_0 = head = feat[0], t_i = feat[1 + 2i], x_{i+1} = feat[2 + 2i], u_0 last);
        # inference needs three buffers: head (read again by body.conv's residual), x and t.  x_{i+1} overwrites x_i
        # in place for i >= 1: conv2 reads t, and x_i only enters as the residual of the very pixel a thread stores.
        cur = 0                                              # index of x_i in feat
        for i in range(R):
            t = feat[1 + 2 * i] if g.train else feat[2]
            nxt = feat[2 + 2 * i] if g.train else feat[1]
            self._conv(feat[cur], by[f'body.{i}.body.conv1'], t, relu=1)
            self._conv(t, by[f'body.{i}.body.conv2'], nxt, res=feat[cur], scale=self.res_scale)
            cur = 2 + 2 * i if g.train else 1
        u0 = feat[2 * R + 1] if g.train else feat[2]
        self._conv(feat[cur], by['body.conv'], u0, res=feat[0])
        g.u0 = u0
        src = u0
        for k, r in enumerate(self.factors):
            l = by[f'tail.0.conv{k + 1}']
            bn = 256 if r == 2 else 192
            ops.conv3x3(src, [0], g.n, l.w, bn, epi=L.EPI_PS, kb_per_src=l.fwd.kb_per_src, bias=l.b,
                        n_tiles_n=l.n_total // bn, out_bf16=g.up[k], ps_r=r)
            src = g.up[k]
        l = by['tail.conv']
        ops.conv3x3(src, [0], g.n, l.w, 64, epi=L.EPI_STORE, kb_per_src=l.fwd.kb_per_src, bias=l.b, n_tiles_n=1,
                    out_f32=g.out16, out_ch=16, n_store=16)
        H, W = g.sizes[-1]
        L.check(lib.pvsr_take_channel0_f32(L.ptr(g.out16), 16, L.ptr(g.out), g.n * H * W, st), 'take_channel0')

    def _wgrad_desc(self, g, d, layer, views, src_img_base, dys, out_hw):
        """Fills one pvsr_wgrad_desc: X = view 0 at image offset src_img_base, dY chunks `dys` (view, img_base, ch0,
        off_x, off_y) in packed-column order; result slab of `layer` inside g.dw."""
        d.H, d.W = out_hw
        d.n_img = g.n
        d.n_views = len(views)
        for i, (t, mul) in enumerate(views):
            d.views[i].ptr = t.data_ptr()
            d.views[i].channels = t.shape[3]
            d.views[i].H, d.views[i].W = t.shape[1], t.shape[2]
            d.views[i].images = t.shape[0]
            d.views[i].mul = mul
        d.n_src = 1
        d.src_view[0], d.src_img_base[0] = 0, src_img_base
        d.n_dy = len(dys)
        for i, (v, base, ch0, ox, oy) in enumerate(dys):
            d.dy_view[i], d.dy_img_base[i], d.dy_ch0[i], d.dy_off_x[i], d.dy_off_y[i] = v, base, ch0, ox, oy
        d.kb_per_src, d.taps, d.n_total, d.with_bias, d.n_splits = layer.fwd.kb_per_src, 9, layer.n_total, 1, 0
        off = g.dw_off[layer.name]
        n_w = layer.n_kb * layer.n_total * 64
        d.dw_packed = g.dw.data_ptr() + 4 * off
        d.db_packed = g.dw.data_ptr() + 4 * (off + n_w)

    def _build_wgrad_launches(self, g):
        """Weight-gradient launches of a training geometry: tail, every up-sampler conv, ALL body convs in one launch
        (layer l: X = acts slot l, dY = gacts slot l + 1), head.  Job lists are staged on the device once."""
        lib, st, by, n = L.load(), L.current_stream(), self.by_name, g.n
        sb = lib.pvsr_wgrad_scratch_bytes()
        launches = []

        def add(descs):
            arr = (L.WgradDesc * len(descs))(*descs)
            scratch = torch.empty(sb, dtype=torch.uint8, device=self.device)
            arr[0].job_scratch = scratch.data_ptr()
            L.check(lib.pvsr_conv3x3_wgrad_multi(C.cast(arr, C.c_void_p), len(descs), 1, st), 'wgrad job upload')
            launches.append((arr, len(descs), scratch))

        H, W = g.sizes[-1]
        d = L.WgradDesc()
        self._wgrad_desc(g, d, by['tail.conv'], [(g.up[-1], 1), (g.g64, 1)], 0, [(1, 0, 0, 0, 0)], (H, W))
        add([d])
        for k, r in enumerate(self.factors):
            x_in = g.up[k - 1] if k > 0 else g.feat[2 * self.R + 1]
            d = L.WgradDesc()
            self._wgrad_desc(g, d, by[f'tail.0.conv{k + 1}'], [(x_in, 1), (g.dup[k], r)], 0, self._chunks(1, r),
                             g.sizes[k])
            add([d])
        body = [l for l in self.layers if l.kind == 'body']          # conv1_0, conv2_0, ..., body.conv: slot l -> l + 1
        views = [(g.acts, 1), (g.gacts, 1)]
        for c0 in range(0, len(body), 128):
            descs = []
            for li in range(c0, min(c0 + 128, len(body))):
                d = L.WgradDesc()
                self._wgrad_desc(g, d, body[li], views, li * n,
                                 [(1, (li + 1) * n, c * 64, 0, 0) for c in range(self.F // 64)], (g.h, g.w))
                descs.append(d)
            add(descs)
        d = L.WgradDesc()
        self._wgrad_desc(g, d, by['head.0'], [(g.x64, 1), (g.gfeat[0], 1)], 0, self._chunks(1), (g.h, g.w))
        add([d])
        g.wg_launches = launches

    def _build_scatter_table(self, g, grads):
        jobs = []
        for l in self.layers:
            off = g.dw_off[l.name]
            n_w = l.n_kb * l.n_total * 64
            s = self.res_scale if l.name.endswith('.body.conv2') else 1.0      # x_{i+1} = s * conv2(t_i) + x_i
            base = g.dw.data_ptr()
            jobs.append(L.TableJob(base + 4 * off, l.idx_w.data_ptr(), grads[l.name + '.weight'].data_ptr(), n_w, s,
                                   L.TJ_SCATTER))
            jobs.append(L.TableJob(base + 4 * (off + n_w), l.idx_b.data_ptr(), grads[l.name + '.bias'].data_ptr(),
                                   l.n_total, s, L.TJ_SCATTER))
        g.scatter_table, g.scatter_max = self._upload_table(jobs)
        g.scatter_jobs = len(jobs)

    def _dgrad(self, src, layer, out, mask=None, res=None, scale=0.0, views=None, srcs=None, kb=None, k16_last=4):
        bn, nt = _tile(self.F)
        ops.conv3x3(views if views is not None else src, srcs if srcs is not None else [0], out.shape[0], layer.wt, bn,
                    epi=L.EPI_STORE, kb_per_src=kb if kb is not None else layer.bwd.kb_per_src, k16_last=k16_last,
                    n_tiles_n=nt, out_bf16=out, res=res, mask=mask, out_scale=scale,
                    out_hw=(out.shape[1], out.shape[2]))

    def _chunks(self, view, r=1):
        """dY chunk list of an F-channel gradient tensor: plain (r = 1) or per sub-pixel of a pixel-unshuffled view."""
        cb = self.F // 64
        if r == 1:
            return [(view, 0, c * 64, 0, 0) for c in range(cb)]
        return [(view, 0, c * 64, q % r, q // r) for q in range(r * r) for c in range(cb)]

    def _backward_launches(self, g):
        """Gradients of everything wrt g.dout -> packed weight gradients in g.dw -> the scatter table's targets."""
        lib, st = L.load(), L.current_stream()
        R, by, s = self.R, self.by_name, self.res_scale
        H, W = g.sizes[-1]
        gf = g.gfeat
        g.dw.zero_()
        L.check(lib.pvsr_pad_channel_bf16(L.ptr(g.dout), L.ptr(g.g64), g.n * H * W, st), 'pad_channel(dout)')
        # ---- data gradients, output to input
        self._dgrad(g.g64, by['tail.conv'], g.dup[-1], kb=1, k16_last=1)
        for k in reversed(range(len(self.factors))):
            r = self.factors[k]
            d_in = g.dup[k - 1] if k > 0 else gf[2 * R + 1]
            self._dgrad(None, by[f'tail.0.conv{k + 1}'], d_in, views=[(g.dup[k], r)],
                        srcs=[(0, 0, 0, q % r, q // r) for q in range(r * r)])
        self._dgrad(gf[2 * R + 1], by['body.conv'], gf[2 * R])                 # u_0 = conv(x_R) + head
        for i in reversed(range(R)):
            # x_{i+1} = s * conv2(t_i) + x_i;  t_i = relu(conv1(x_i))
            self._dgrad(gf[2 + 2 * i], by[f'body.{i}.body.conv2'], gf[1 + 2 * i], mask=g.feat[1 + 2 * i], scale=s)
            self._dgrad(gf[1 + 2 * i], by[f'body.{i}.body.conv1'], gf[2 * i], res=gf[2 + 2 * i])
        # the head output feeds block 0 (gradient now in slot 0) and body.conv's residual (slot 2R + 1)
        L.check(lib.pvsr_add_bf16(L.ptr(gf[0]), L.ptr(gf[2 * R + 1]), L.ptr(gf[0]), gf[0].numel(), st), 'add_bf16')
        # ---- weight gradients
        for arr, n_desc, _ in g.wg_launches:
            L.check(lib.pvsr_conv3x3_wgrad_multi(C.cast(arr, C.c_void_p), n_desc, 0, st), 'wgrad')
        L.check(lib.pvsr_run_table(L.ptr(g.scatter_table), g.scatter_jobs, g.scatter_max, st), 'scatter table')

    # ------------------------------------------------------------------------------------------ public
    def _replay(self, g, which, fn):
        """Runs `fn` eagerly the first two times (warm-up: lazy attribute setup inside the library must not happen
        under capture), then captures it into a CUDA graph and replays."""
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

    def forward(self, x, train=False, clone=True):
        g = self.geometry(x, train)
        self._ensure_packed()
        g.x32.copy_(x.reshape(g.n, g.h, g.w))
        self._replay(g, 'fwd', lambda: self._forward_launches(g))
        return (g.out.clone() if clone else g.out), g

    def backward(self, g, grads):
        """Backward of the last forward(train=True) of geometry `g`; g.dout holds dL/d(out)."""
        if not g.jobs_ready:
            self._build_wgrad_launches(g)               # stages the wgrad job lists on the device
            g.jobs_ready = True
        key = tuple(grads[k].data_ptr() for k in sorted(grads))
        if getattr(g, 'bwd_key', None) != key:          # a captured graph writes to the buffers it was captured with
            self._build_scatter_table(g, grads)
            g.bwd_key, g.graph_bwd = key, None
        self._replay(g, 'bwd', lambda: self._backward_launches(g))

    def grad_buffers(self):
        if getattr(self, '_grad_buf', None) is None:
            self._grad_buf = {k: torch.zeros_like(p) for k, p in self._named().items()}
        return self._grad_buf

    def loss_and_grads(self, x, target, zero_grads=True):
        """Fused training-step body: forward, nn.L1Loss(output, target) (acdc_sisr_trainer.py:27-37 with the L1Loss
        of configs/train/edsr_net/exp1_x4.yaml:44-46) and backward into `p.grad`.  Returns (loss, output)."""
        out, g = self.forward(x, train=True, clone=False)
        g.target.copy_(target.reshape(g.target.shape))
        n = g.out.numel()
        key = ('lw', n)
        w = getattr(self, '_lw', {}).get(key)
        if w is None:
            self._lw = getattr(self, '_lw', {})
            w = self._lw[key] = torch.tensor([1.0 / n], dtype=torch.float32, device=self.device)
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
        return g.loss.clone(), out


class _EDSRFunction(torch.autograd.Function):
    """Autograd bridge: `net(input)` in training mode is differentiable wrt the parameters, so the reference's SISR
    trainer sequence (acdc_sisr_trainer.py / base_trainer.py: any torch loss, loss.backward(), any torch optimiser)
    works unchanged; the input image receives no gradient (the reference never asks for one)."""

    @staticmethod
    def forward(ctx, engine, x, names, *params):
        out, g = engine.forward(x, train=True, clone=True)
        g.fwd_serial = getattr(g, 'fwd_serial', 0) + 1
        ctx.engine, ctx.g, ctx.names, ctx.serial = engine, g, names, g.fwd_serial
        return out

    @staticmethod
    def backward(ctx, grad_out):
        engine, g = ctx.engine, ctx.g
        if g.fwd_serial != ctx.serial:
            raise L.PvsrError('the activations saved by this forward were overwritten by a later forward of the same shape (the plan keeps ONE set of training buffers per shape): call backward before the next forward')
        g.dout.copy_(grad_out.reshape(g.dout.shape))
        bufs = engine.grad_buffers()
        for b in bufs.values():
            b.zero_()
        engine.backward(g, bufs)
        return (None, None, None) + tuple(bufs[k].clone() for k in ctx.names)


def edsr_train_forward(net, x):
    named = list(net.named_parameters())
    return _EDSRFunction.apply(net.engine, x.detach(), tuple(k for k, _ in named), *[p for _, p in named])
