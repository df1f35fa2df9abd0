"""Per-op Python wrappers over the C ABI (torch tensors in, torch tensors out).

Used by the unit tests and by tools; the network itself runs through the plan API (pvsr/engine.py).
Layouts: activations are bf16 NHWC stacked frame-major as [images, H, W, C].
"""
import ctypes as C

import numpy as np
import torch

from . import lib as L


# ------------------------------------------------------------------------------------------------ pack specs
def _spec(c_out, c_in, k, n_src, offs, src_ch, kb_per_src, taps, n_total, ps_r=0, transpose_flip=0, k_ps_r=0):
    s = L.PackSpec()
    s.c_out, s.c_in, s.kh, s.kw = c_out, c_in, k, k
    s.n_src = n_src
    for i, o in enumerate(offs):
        s.src_ch_off[i] = o
    s.src_ch, s.kb_per_src, s.taps, s.n_total = src_ch, kb_per_src, taps, n_total
    s.ps_r, s.transpose_flip, s.k_ps_r = ps_r, transpose_flip, k_ps_r
    return s


def spec_lstm(feat=64):
    """ConvLSTMCell.conv (refine_net.py:235): in = [x | h], out = [i | f | o | g]."""
    return _spec(4 * feat, 2 * feat, 3, 2, [0, feat], feat, 1, 9, 4 * feat)


def spec_refine_conv1(window=5, feat=64):
    """_RefineBlock.body.conv1 with positional encoding (refine_net.py:149): per frame [fwd 64 | bwd 64 | pos 1].
    Sources are ordered (frame 0 fwd, frame 0 bwd, frame 1 fwd, ...); the pos channel is handled by posterm."""
    per = 2 * feat + 1
    offs = [per * (s // 2) + feat * (s % 2) for s in range(2 * window)]
    return _spec(per, per * window, 3, 2 * window, offs, feat, 1, 9, 144)


def spec_refine_conv2(feat=64):
    """_RefineBlock.body.conv2 (refine_net.py:151): 129 -> 64; the 129-channel input is stored with 144 channels."""
    return _spec(feat, 2 * feat + 1, 3, 1, [0], 2 * feat + 1, 3, 9, feat)


def spec_refine_conv1x1(window=5, feat=64):
    """_RefineBlock.body.conv1 without positional encoding (refine_net.py:154): 1x1 conv over [fwd | bwd] x window."""
    offs = [2 * feat * (s // 2) + feat * (s % 2) for s in range(2 * window)]
    return _spec(feat, 2 * feat * window, 1, 2 * window, offs, feat, 1, 1, feat)


def spec_head_ps(r, feat=64):
    """_OutBlock conv + PixelShuffle(r) (refine_net.py:199-204): columns regrouped per sub-pixel."""
    return _spec(feat * r * r, feat, 3, 1, [0], feat, 1, 9, feat * r * r, ps_r=r)


def pack_index(spec):
    lib = L.load()
    n = lib.pvsr_pack_index_count(C.byref(spec))
    idx = np.empty(n, dtype=np.int32)
    L.check(lib.pvsr_pack_index_host(C.byref(spec), idx.ctypes.data_as(C.c_void_p)), "pack_index")
    return idx


def pack_bias_index(spec):
    lib = L.load()
    idx = np.empty(spec.n_total, dtype=np.int32)
    L.check(lib.pvsr_pack_bias_index_host(C.byref(spec), idx.ctypes.data_as(C.c_void_p)), "pack_bias_index")
    return idx


def pack_weight(weight, spec, weight_sum_spec=None):
    """fp32 Conv2d weight (cuda) -> packed bf16 [rows, 64]."""
    lib = L.load()
    idx = torch.from_numpy(pack_index(spec)).to(weight.device)
    idx2 = torch.from_numpy(pack_index(weight_sum_spec)).to(weight.device) if weight_sum_spec is not None else None
    out = torch.empty(idx.numel() // 64, 64, dtype=torch.bfloat16, device=weight.device)
    w = weight.detach().contiguous().float()
    L.check(lib.pvsr_pack_weights(L.ptr(w), L.ptr(idx), L.ptr(idx2), L.ptr(out), idx.numel(), L.current_stream()),
            "pack_weights")
    return out


def pack_bias(bias, spec):
    lib = L.load()
    idx = torch.from_numpy(pack_bias_index(spec)).to(bias.device)
    out = torch.empty(spec.n_total, dtype=torch.float32, device=bias.device)
    b = bias.detach().contiguous().float()
    L.check(lib.pvsr_gather_f32(L.ptr(b), L.ptr(idx), L.ptr(out), idx.numel(), L.current_stream()), "gather_f32")
    return out


# ------------------------------------------------------------------------------------------------ ops
def in_conv_prelu(x, w, b, slope):
    """x fp32 [n_img, H, W] -> bf16 [n_img, H, W, 64]   (_InBlock, refine_net.py:188-192)."""
    lib = L.load()
    n, H, W = x.shape
    out = torch.empty(n, H, W, 64, dtype=torch.bfloat16, device=x.device)
    L.check(lib.pvsr_in_conv_prelu_fwd(L.ptr(x.contiguous()), L.ptr(w.contiguous()), L.ptr(b.contiguous()),
                                       L.ptr(slope.contiguous()), L.ptr(out), n, H, W, L.current_stream()),
            "in_conv_prelu")
    return out


def conv3x3(act, srcs, n_img, w_packed, bn, epi=L.EPI_STORE, kb_per_src=1, k16_last=4, taps=9, bias=None,
            n_tiles_n=1, w_row_base=0, out_bf16=None, out_f32=None, res=None, posterm=None, out_ch=None,
            n_store=None, ps_r=0, c_in=None, c_out=None, h_out=None, gates_out=None, grad0=None, grad1=None,
            grad_split=0, out_hw=None, relu=0, mask=None, out_scale=0.0, prelu=None):
    """Generic launch of the tcgen05 implicit-GEMM conv.

    act  : one bf16 tensor [images, H, W, C], or a list of views (tensor, mul) - mul > 1 reads the pixel-unshuffled
           image of a tensor mul x larger than the output.
    srcs : per source either an image base (view 0) or a tuple (view, img_base, ch0, off_x, off_y).
    """
    lib = L.load()
    views = act if isinstance(act, (list, tuple)) else [(act, 1)]
    d = L.ConvDesc()
    d.epi, d.bn = epi, bn
    if out_hw is None:
        t0, m0 = views[0]
        out_hw = (t0.shape[1] // m0, t0.shape[2] // m0)
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
    for i, v in enumerate(srcs):
        view, base, ch0, ox, oy = v if isinstance(v, (list, tuple)) else (0, v, 0, 0, 0)
        d.src_view[i], d.src_img_base[i], d.src_ch0[i], d.src_off_x[i], d.src_off_y[i] = view, base, ch0, ox, oy
    d.kb_per_src, d.k16_last, d.taps = kb_per_src, k16_last, taps
    d.w_packed = w_packed.data_ptr()
    d.w_rows = w_packed.shape[0]
    d.w_row_base = w_row_base
    d.n_tiles_n = n_tiles_n
    for name, t in (("bias", bias), ("out_bf16", out_bf16), ("out_f32", out_f32), ("res", res),
                    ("posterm", posterm), ("c_in", c_in), ("c_out", c_out), ("h_out", h_out),
                    ("gates_out", gates_out), ("grad0", grad0), ("grad1", grad1), ("mask", mask), ("prelu", prelu)):
        setattr(d, name, None if t is None else t.data_ptr())
    d.out_ch = out_ch if out_ch is not None else bn * n_tiles_n
    d.n_store = n_store if n_store is not None else bn
    d.ps_r = ps_r
    d.grad_split = grad_split
    d.relu, d.out_scale = int(relu), float(out_scale)
    L.check(lib.pvsr_conv3x3_fwd(C.byref(d), L.current_stream()), "conv3x3")


def lstm_state(n_img, H, W, device):
    n = L.load().pvsr_lstm_state_elems(n_img, H, W)
    return torch.zeros(n, dtype=torch.float32, device=device)


def _lstm_tile_index(H, W, device):
    """(wp, tiles_per_img, flat index [H*W] of pixel (y, x) inside an image's [tiles][128] row space or None)."""
    lib = L.load()
    wp, tiles = C.c_int(), C.c_int()
    L.check(lib.pvsr_lstm_tile_geometry(H, W, C.byref(wp), C.byref(tiles)), "lstm_tile_geometry")
    if wp.value == 0:
        return 0, tiles.value, None
    y = torch.arange(H, device=device).view(H, 1)
    x = torch.arange(W, device=device).view(1, W)
    return wp.value, tiles.value, (y * wp.value + x).reshape(-1)


def lstm_state_to_nchw(state, n_img, H, W):
    """Tile-transposed cell state -> [n_img, 64, H, W] fp32 (test helper; layout: pvsr_lstm_tile_geometry)."""
    lib = L.load()
    wp, tiles, pos = _lstm_tile_index(H, W, state.device)
    if wp:
        s = state[:n_img * tiles * 64 * 128].view(n_img, tiles, 64, 128).permute(0, 2, 1, 3).reshape(n_img, 64, tiles * 128)
        return s[:, :, pos].reshape(n_img, 64, H, W).contiguous()
    l = C.c_int()
    lib.pvsr_choose_tile(H, W, C.byref(l))
    tw, th = 1 << l.value, 128 >> l.value
    tx, ty = (W + tw - 1) // tw, (H + th - 1) // th
    s = state[:n_img * ty * tx * 64 * 128].view(n_img, ty, tx, 64, th, tw).permute(0, 3, 1, 4, 2, 5)
    return s.reshape(n_img, 64, ty * th, tx * tw)[:, :, :H, :W].contiguous()


def refine_posterm(w1, b1, pos, n_frames_out, window=5, feat=64, n_total=144):
    """pos fp32 [B, L] -> table fp32 [n_frames_out * B, 16, n_total]."""
    lib = L.load()
    B, Lf = pos.shape
    c_out, c_in = w1.shape[0], w1.shape[1]
    table = torch.empty(n_frames_out * B, 16, n_total, dtype=torch.float32, device=pos.device)
    L.check(lib.pvsr_refine_posterm(L.ptr(w1.contiguous()), L.ptr(b1.contiguous()), L.ptr(pos.contiguous()),
                                    L.ptr(table), n_frames_out, B, Lf, window, c_out, c_in, 2 * feat, n_total,
                                    L.current_stream()), "posterm")
    return table


def head_conv_last(x, w, b, target=None, l1_partial=None):
    """bf16 [n_img, H, W, 64] -> fp32 [n_img, H, W]   (_OutBlock last conv, refine_net.py:203/205)."""
    lib = L.load()
    n, H, W, _ = x.shape
    out = torch.empty(n, H, W, dtype=torch.float32, device=x.device)
    L.check(lib.pvsr_head_conv_last_fwd(L.ptr(x), L.ptr(w.contiguous()), L.ptr(b.contiguous()), L.ptr(out),
                                        L.ptr(target), L.ptr(l1_partial), n, H, W, L.current_stream()),
            "head_conv_last")
    return out


def add_bf16(a, b):
    lib = L.load()
    out = torch.empty_like(a)
    L.check(lib.pvsr_add_bf16(L.ptr(a), L.ptr(b), L.ptr(out), a.numel(), L.current_stream()), "add_bf16")
    return out


def conv3x3_wgrad(views, srcs, dys, n_img, out_hw, n_total, kb_per_src=1, taps=9, with_bias=True, n_splits=0,
                  dw=None, db=None):
    """Weight/bias gradient in packed layout (fp32, accumulated).
    views: list of (tensor [images,H,W,C] bf16, mul); srcs / dys: tuples (view, img_base, ch0, off_x, off_y)."""
    lib = L.load()
    dev = views[0][0].device
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
    d.kb_per_src, d.taps, d.n_total, d.with_bias, d.n_splits = kb_per_src, taps, n_total, int(with_bias), n_splits
    n_kb = len(srcs) * taps * kb_per_src
    buf = torch.zeros(n_kb * n_total * 64 + n_total, dtype=torch.float32, device=dev) if dw is None else None
    if dw is None:
        dw, db = buf[:n_kb * n_total * 64], buf[n_kb * n_total * 64:]
    d.dw_packed, d.db_packed = dw.data_ptr(), db.data_ptr()
    scratch = torch.empty(lib.pvsr_wgrad_scratch_bytes(), dtype=torch.uint8, device=dev)
    d.job_scratch = scratch.data_ptr()
    L.check(lib.pvsr_conv3x3_wgrad(C.byref(d), L.current_stream()), "conv3x3_wgrad")
    return dw.view(n_kb, n_total, 64), db


def scatter_add(param_grad, spec, packed, bias_grad=None, packed_bias=None):
    """Packed fp32 gradient -> parameter-layout gradient (+=) through the packing index."""
    lib = L.load()
    idx = torch.from_numpy(pack_index(spec)).to(packed.device)
    L.check(lib.pvsr_scatter_add(L.ptr(param_grad), L.ptr(idx), None, L.ptr(packed), idx.numel(), L.current_stream()),
            "scatter_add")
    if bias_grad is not None:
        bidx = torch.from_numpy(pack_bias_index(spec)).to(packed.device)
        L.check(lib.pvsr_scatter_add(L.ptr(bias_grad), L.ptr(bidx), None, L.ptr(packed_bias), bidx.numel(),
                                     L.current_stream()), "scatter_add")


# ------------------------------------------------------------------------------------------------ backward ops
def lstm_cell_bwd_pointwise(dh, gates, c, c_prev, dc, dc_zero, n_img, H, W):
    """dh fp32 [n_img,H,W,64]; gates bf16 / c, c_prev, dc fp32 tile-transposed -> dgates bf16 [n_img,H,W,256]."""
    lib = L.load()
    dg = torch.zeros(n_img, H, W, 256, dtype=torch.bfloat16, device=dh.device)
    L.check(lib.pvsr_lstm_cell_bwd_pointwise(L.ptr(dh), L.ptr(gates), L.ptr(c), L.ptr(c_prev), L.ptr(dc), int(dc_zero),
                                             L.ptr(dg), n_img, H, W, L.current_stream()), "lstm_cell_bwd_pointwise")
    return dg


def nchw_to_lstm_state(x, dtype=torch.float32):
    """[n_img, C, H, W] -> tile-transposed [tile][C][128] buffer (test helper, inverse of lstm_state_to_nchw)."""
    lib = L.load()
    n, Cc, H, W = x.shape
    wp, tiles, pos = _lstm_tile_index(H, W, x.device)
    if wp:
        flat = torch.zeros(n, Cc, tiles * 128, dtype=x.dtype, device=x.device)
        flat[:, :, pos] = x.reshape(n, Cc, H * W)
        return flat.view(n, Cc, tiles, 128).permute(0, 2, 1, 3).contiguous().reshape(-1).to(dtype)
    l = C.c_int()
    lib.pvsr_choose_tile(H, W, C.byref(l))
    tw, th = 1 << l.value, 128 >> l.value
    tx, ty = (W + tw - 1) // tw, (H + th - 1) // th
    pad = torch.zeros(n, Cc, ty * th, tx * tw, dtype=x.dtype, device=x.device)
    pad[:, :, :H, :W] = x
    s = pad.view(n, Cc, ty, th, tx, tw).permute(0, 2, 4, 1, 3, 5).contiguous()
    return s.reshape(-1).to(dtype)


def l1_multistage(out, target, weights, want_grad=True):
    """out [lists, ...], target [...], weights [lists] -> (loss 0-dim, dout like out or None)."""
    lib = L.load()
    n_lists = out.shape[0]
    n_per = target.numel()
    loss = torch.zeros((), dtype=torch.float32, device=out.device)
    dout = torch.empty_like(out) if want_grad else None
    L.check(lib.pvsr_l1_multistage(L.ptr(out), L.ptr(target), L.ptr(weights), n_lists, n_per, L.ptr(loss), L.ptr(dout),
                                   L.current_stream()), "l1_multistage")
    return loss, dout


def head_conv_last_bwd(x, w, dout):
    """x bf16 [n,H,W,64], w (1,64,3,3), dout fp32 [n,H,W] -> (din bf16 [n,H,W,64], dw (1,64,3,3), db (1))."""
    lib = L.load()
    n, H, W, _ = x.shape
    din = torch.empty_like(x)
    dw = torch.zeros(1, 64, 3, 3, dtype=torch.float32, device=x.device)
    db = torch.zeros(1, dtype=torch.float32, device=x.device)
    L.check(lib.pvsr_head_conv_last_bwd_data(L.ptr(dout), L.ptr(w.contiguous()), L.ptr(din), n, H, W,
                                             L.current_stream()), "head_conv_last_bwd_data")
    L.check(lib.pvsr_head_conv_last_bwd_weight(L.ptr(x), L.ptr(dout), L.ptr(dw), L.ptr(db), n, H, W,
                                               L.current_stream()), "head_conv_last_bwd_weight")
    return din, dw, db


def head_tail_bwd(x, w2, b2, w3, dout, sign_scale=0.0):
    """Rank-1 adjoint of conv3x3 (64 -> 256) + PixelShuffle(2) + conv3x3 (64 -> 1) (csrc/tail_rank1.cu).
    x bf16 [n,H1,W1,64], w2 (256,64,3,3), b2 (256), w3 (1,64,3,3), dout fp32 [n,2 H1,2 W1]
    -> (dx bf16 [n,H1,W1,64], dw2, db2, dw3, db3)."""
    lib = L.load()
    n, H1, W1, _ = x.shape
    dx = torch.empty_like(x)
    dw2, db2 = torch.zeros_like(w2), torch.zeros_like(b2)
    dw3 = torch.zeros_like(w3)
    db3 = torch.zeros(1, dtype=torch.float32, device=x.device)
    scratch = torch.empty(lib.pvsr_head_tail_scratch_bytes(), dtype=torch.uint8, device=x.device)
    L.check(lib.pvsr_head_tail_bwd(L.ptr(dout.contiguous()), L.ptr(x), L.ptr(w2.contiguous()), L.ptr(b2.contiguous()),
                                   L.ptr(w3.contiguous()), L.ptr(dx), L.ptr(dw2), L.ptr(db2), L.ptr(dw3), L.ptr(db3),
                                   L.ptr(scratch), n, H1, W1, float(sign_scale), L.current_stream()), "head_tail_bwd")
    return dx, dw2, db2, dw3, db3


def head_tail_fwd(x, w2, b2, w3, b3):
    """Composite forward of conv3x3 (64 -> 256) + PixelShuffle(2) + conv3x3 (64 -> 1) (csrc/tail_rank1.cu).
    x bf16 [n,H1,W1,64] -> out fp32 [n, 2 H1, 2 W1]."""
    lib = L.load()
    n, H1, W1, _ = x.shape
    tables = torch.empty(lib.pvsr_head_tail_fwd_table_bytes(), dtype=torch.uint8, device=x.device)
    L.check(lib.pvsr_head_tail_fwd_tables(L.ptr(w2.contiguous()), L.ptr(b2.contiguous()), L.ptr(w3.contiguous()),
                                          L.ptr(b3.contiguous()), L.ptr(tables), L.current_stream()), "head_tail_fwd_tables")
    out = torch.empty(n, 2 * H1, 2 * W1, dtype=torch.float32, device=x.device)
    L.check(lib.pvsr_head_tail_fwd(L.ptr(x), L.ptr(tables), L.ptr(out), n, H1, W1, L.current_stream()), "head_tail_fwd")
    return out


def in_conv_prelu_bwd(x, w, b, slope, g):
    """x fp32 [n,H,W], g fp32 [n,H,W,64] -> (dw (64,1,3,3), db (64), dslope (1))."""
    lib = L.load()
    n, H, W = x.shape
    dw = torch.zeros(64, 1, 3, 3, dtype=torch.float32, device=x.device)
    db = torch.zeros(64, dtype=torch.float32, device=x.device)
    da = torch.zeros(1, dtype=torch.float32, device=x.device)
    L.check(lib.pvsr_in_conv_prelu_bwd(L.ptr(x.contiguous()), L.ptr(w.contiguous()), L.ptr(b.contiguous()),
                                       L.ptr(slope.contiguous()), L.ptr(g.contiguous()), L.ptr(dw), L.ptr(db),
                                       L.ptr(da), n, H, W, L.current_stream()), "in_conv_prelu_bwd")
    return dw, db, da


def refine_posterm_bwd(g, pos, dw1, n_frames, frame0, window=5, feat=64):
    """g bf16 [n_frames*B, H, W, ch]; pos fp32 [B, L]; dw1 (c_out, c_in, 3, 3) += on the pos channels."""
    lib = L.load()
    B, Lf = pos.shape
    _, H, W, ch = g.shape
    sums = torch.empty(window, 16, ch, dtype=torch.float32, device=g.device)
    L.check(lib.pvsr_refine_posterm_bwd(L.ptr(g), L.ptr(pos.contiguous()), L.ptr(sums), L.ptr(dw1), n_frames, B, Lf,
                                        frame0, window, H, W, dw1.shape[0], dw1.shape[1], 2 * feat, ch,
                                        L.current_stream()), "refine_posterm_bwd")
    return sums


def cast_f32_bf16(x):
    lib = L.load()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    L.check(lib.pvsr_cast_f32_bf16(L.ptr(x.contiguous()), L.ptr(out), x.numel(), L.current_stream()), "cast_f32_bf16")
    return out
