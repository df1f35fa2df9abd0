"""ctypes binding of libpvsr.so (include/pvsr.h).

The product path is CUDA only: if the shared library is missing this module raises at import of the
symbols (no CPU fallback, no oracle on this path).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")
LIB_PATH = os.path.join(CSRC, "libpvsr.so")

MAX_SRC = 10
MAX_LAYERS = 8
MAX_HEAD_CONVS = 4
MAX_VIEWS = 4
EPI_STORE, EPI_PS, EPI_LSTM, EPI_GRAD = 0, 1, 2, 3

c_void_p, c_int, c_int64, c_float_p, c_int32_p = C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p


class PackSpec(C.Structure):
    _fields_ = [("c_out", c_int), ("c_in", c_int), ("kh", c_int), ("kw", c_int), ("n_src", c_int),
                ("src_ch_off", c_int * MAX_SRC), ("src_ch", c_int), ("kb_per_src", c_int), ("taps", c_int),
                ("n_total", c_int), ("ps_r", c_int), ("transpose_flip", c_int), ("k_ps_r", c_int),
                ("src_col_off", c_int * MAX_SRC), ("ps_ch", c_int)]


class ActView(C.Structure):
    _fields_ = [("ptr", c_void_p), ("channels", c_int), ("W", c_int), ("H", c_int), ("images", c_int64),
                ("mul", c_int)]


class ConvDesc(C.Structure):
    _fields_ = [("epi", c_int), ("bn", c_int), ("H", c_int), ("W", c_int), ("n_img", c_int64),
                ("n_views", c_int), ("views", ActView * MAX_VIEWS), ("n_src", c_int),
                ("src_view", c_int * MAX_SRC), ("src_img_base", c_int * MAX_SRC), ("src_ch0", c_int * MAX_SRC),
                ("src_off_x", c_int * MAX_SRC), ("src_off_y", c_int * MAX_SRC),
                ("kb_per_src", c_int), ("k16_last", c_int), ("taps", c_int),
                ("w_packed", c_void_p), ("w_rows", c_int64), ("w_row_base", c_int), ("n_tiles_n", c_int),
                ("bias", c_void_p), ("out_bf16", c_void_p), ("out_f32", c_void_p), ("res", c_void_p),
                ("posterm", c_void_p), ("out_ch", c_int), ("n_store", c_int), ("ps_r", c_int),
                ("grad0", c_void_p), ("grad1", c_void_p), ("grad_split", c_int),
                ("c_in", c_void_p), ("c_out", c_void_p), ("h_out", c_void_p), ("gates_out", c_void_p),
                ("relu", c_int), ("mask", c_void_p), ("out_scale", C.c_float), ("prelu", c_void_p)]


MAX_DY = 36


class WgradDesc(C.Structure):
    _fields_ = [("H", c_int), ("W", c_int), ("n_img", c_int64), ("n_views", c_int), ("views", ActView * MAX_VIEWS),
                ("n_src", c_int), ("src_view", c_int * MAX_SRC), ("src_img_base", c_int * MAX_SRC),
                ("src_ch0", c_int * MAX_SRC), ("src_off_x", c_int * MAX_SRC), ("src_off_y", c_int * MAX_SRC),
                ("kb_per_src", c_int), ("taps", c_int), ("n_dy", c_int), ("dy_view", c_int * MAX_DY),
                ("dy_img_base", c_int * MAX_DY), ("dy_ch0", c_int * MAX_DY), ("dy_off_x", c_int * MAX_DY),
                ("dy_off_y", c_int * MAX_DY), ("n_total", c_int), ("with_bias", c_int), ("dw_packed", c_void_p),
                ("db_packed", c_void_p), ("n_splits", c_int), ("job_scratch", c_void_p)]


class NetConfig(C.Structure):
    _fields_ = [("batch", c_int), ("n_frames", c_int), ("n_updated", c_int), ("h", c_int), ("w", c_int),
                ("scale", c_int), ("n_stages", c_int), ("window", c_int), ("n_layers", c_int), ("pos_enc", c_int),
                ("memory", c_int), ("all_heads", c_int), ("save_for_backward", c_int)]


class NetParams(C.Structure):
    _fields_ = [("in_w", c_void_p), ("in_b", c_void_p), ("in_slope", c_void_p),
                ("lstm_w", (c_void_p * MAX_LAYERS) * 2), ("lstm_b", (c_void_p * MAX_LAYERS) * 2),
                ("ref_w1", c_void_p), ("ref_b1", c_void_p), ("ref_w2", c_void_p), ("ref_b2", c_void_p),
                ("head_w", c_void_p * MAX_HEAD_CONVS), ("head_b", c_void_p * MAX_HEAD_CONVS)]


class NetGrads(C.Structure):
    """fp32 gradient buffers in the parameter layouts (accumulated by pvsr_plan_backward)."""
    _fields_ = NetParams._fields_


class CineSample(C.Structure):
    _fields_ = [("vol_off", c_int64), ("pos_off", c_int64), ("T", C.c_int32), ("t_first", C.c_int32),
                ("Hs", C.c_int32), ("Ws", C.c_int32), ("ay", C.c_int32), ("by", C.c_int32), ("ax", C.c_int32),
                ("bx", C.c_int32)]


class TableJob(C.Structure):
    _fields_ = [("src", c_void_p), ("idx", c_void_p), ("dst", c_void_p), ("n", c_int64), ("scale", C.c_float),
                ("kind", c_int)]


TJ_PACK, TJ_GATHER, TJ_SCATTER = 0, 1, 2
DT_F32, DT_I16, DT_U16, DT_U8, DT_F64 = 0, 1, 2, 3, 4
NUM_CLASSES, NUM_CLASSES_BWD = 7, 9

# name -> (restype, argtypes); every symbol declared in include/pvsr.h
SIGNATURES = {
    "pvsr_version": (c_int, []),
    "pvsr_last_error": (C.c_char_p, []),
    "pvsr_device_check": (c_int, []),
    "pvsr_set_cta_pair": (c_int, [c_int]),
    "pvsr_get_cta_pair": (c_int, []),
    "pvsr_set_halo_mode": (c_int, [c_int]),
    "pvsr_get_halo_mode": (c_int, []),
    "pvsr_set_w_resident": (c_int, [c_int]),
    "pvsr_get_w_resident": (c_int, []),
    "pvsr_set_pack_table": (c_int, [c_int]),
    "pvsr_get_pack_table": (c_int, []),
    "pvsr_set_tail_rank1": (c_int, [c_int]),
    "pvsr_get_tail_rank1": (c_int, []),
    "pvsr_head_tail_scratch_bytes": (c_int64, []),
    "pvsr_head_tail_bwd": (c_int, [c_void_p] * 11 + [c_int64, c_int, c_int, C.c_float, c_void_p]),
    "pvsr_debug_dump_trace": (c_int, []),
    "pvsr_debug_clear_trace": (None, []),
    "pvsr_plan_set_sign_gradient": (c_int, [c_void_p, c_void_p, c_int]),
    "pvsr_set_tail_fwd": (c_int, [c_int]),
    "pvsr_get_tail_fwd": (c_int, []),
    "pvsr_head_tail_fwd_table_bytes": (c_int64, []),
    "pvsr_head_tail_fwd_tables": (c_int, [c_void_p] * 6),
    "pvsr_head_tail_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "pvsr_set_two_branch": (c_int, [c_int]),
    "pvsr_get_two_branch": (c_int, []),
    "pvsr_set_pdl": (c_int, [c_int]),
    "pvsr_get_pdl": (c_int, []),
    "pvsr_set_head_tma": (c_int, [c_int]),
    "pvsr_get_head_tma": (c_int, []),
    "pvsr_choose_tile": (c_int, [c_int, c_int, C.POINTER(c_int)]),
    "pvsr_pack_index_count": (c_int64, [C.POINTER(PackSpec)]),
    "pvsr_pack_index_host": (c_int, [C.POINTER(PackSpec), c_void_p]),
    "pvsr_pack_bias_index_host": (c_int, [C.POINTER(PackSpec), c_void_p]),
    "pvsr_pack_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pvsr_gather_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pvsr_in_conv_prelu_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                       c_void_p]),
    "pvsr_conv3x3_fwd": (c_int, [C.POINTER(ConvDesc), c_void_p]),
    "pvsr_scatter_add_scaled": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, C.c_float, c_void_p]),
    "pvsr_lstm_state_elems": (c_int64, [c_int64, c_int, c_int]),
    "pvsr_lstm_tile_geometry": (c_int, [c_int, c_int, C.POINTER(c_int), C.POINTER(c_int)]),
    "pvsr_wgrad_scratch_bytes": (c_int64, []),
    "pvsr_conv3x3_wgrad": (c_int, [C.POINTER(WgradDesc), c_void_p]),
    "pvsr_conv3x3_wgrad_staged": (c_int, [C.POINTER(WgradDesc), c_int, c_void_p]),
    "pvsr_conv3x3_wgrad_multi": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "pvsr_run_table": (c_int, [c_void_p, c_int, c_int64, c_void_p]),
    "pvsr_frame_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, C.c_float, C.c_float, c_void_p,
                                  C.c_float, c_void_p, c_void_p]),
    "pvsr_bicubic_upsample": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "pvsr_pad_channel_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pvsr_take_channel0_f32": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    "pvsr_scatter_add": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pvsr_refine_posterm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_int, c_int, c_void_p]),
    "pvsr_head_conv_last_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                        c_int, c_void_p]),
    "pvsr_cine_gather": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, C.c_double, C.c_double, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "pvsr_add_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pvsr_lstm_cell_bwd_pointwise": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64,
                                             c_int, c_int, c_void_p]),
    "pvsr_l1_multistage": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "pvsr_head_conv_last_bwd_data": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "pvsr_head_conv_last_bwd_weight": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                               c_void_p]),
    "pvsr_in_conv_prelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int64, c_int, c_int, c_void_p]),
    "pvsr_refine_posterm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 11 + [c_void_p]),
    "pvsr_cast_f32_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pvsr_prelu_fwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pvsr_prelu_bwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pvsr_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, C.c_float, C.c_float, C.c_float,
                               C.c_float, C.c_float, C.c_float, c_void_p, c_void_p]),
    "pvsr_plan_backward": (c_int, [c_void_p, C.POINTER(NetParams), c_void_p, c_void_p, c_void_p, c_void_p,
                                   C.POINTER(NetGrads), c_void_p, c_int, c_void_p]),
    "pvsr_plan_num_launches_bwd": (c_int64, [c_void_p]),
    "pvsr_plan_flops_bwd": (C.c_double, [c_void_p]),
    "pvsr_plan_class_stats_bwd": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pvsr_plan_profile_bwd": (c_int, [c_void_p, C.POINTER(NetParams), c_void_p, c_void_p, c_void_p, c_void_p,
                                      C.POINTER(NetGrads), c_void_p, c_void_p, c_void_p]),
    "pvsr_plan_create": (c_int, [C.POINTER(NetConfig), C.POINTER(c_void_p)]),
    "pvsr_plan_destroy": (None, [c_void_p]),
    "pvsr_plan_workspace_bytes": (c_int64, [c_void_p]),
    "pvsr_plan_packed_bytes": (c_int64, [c_void_p]),
    "pvsr_plan_output_elems": (c_int64, [c_void_p]),
    "pvsr_plan_num_lists": (c_int, [c_void_p]),
    "pvsr_plan_num_launches": (c_int64, [c_void_p]),
    "pvsr_plan_flops": (C.c_double, [c_void_p]),
    "pvsr_plan_pack": (c_int, [c_void_p, C.POINTER(NetParams), c_void_p, c_void_p]),
    "pvsr_plan_forward": (c_int, [c_void_p, C.POINTER(NetParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int, c_void_p]),
    "pvsr_plan_class_stats": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pvsr_plan_profile": (c_int, [c_void_p, C.POINTER(NetParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
}

_lib = None


class PvsrError(RuntimeError):
    pass


def load():
    """Loads libpvsr.so (once). Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PvsrError(f"{LIB_PATH} is missing - build it with `python {os.path.join(CSRC, 'build.py')}` "
                        "(or __graft_entry__.build()); the CUDA extension is the only compute path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if os.environ.get("PVSR_HALO") is not None:           # A/B switch of the halo (slab) conv kernel
        lib.pvsr_set_halo_mode(int(os.environ["PVSR_HALO"]))
    if os.environ.get("PVSR_HEAD_TMA") is not None:       # A/B switch of the head_conv_last forms
        lib.pvsr_set_head_tma(int(os.environ["PVSR_HEAD_TMA"]))
    if os.environ.get("PVSR_W_RESIDENT") is not None:     # A/B switch of the resident weight operand (narrow slab launches)
        lib.pvsr_set_w_resident(int(os.environ["PVSR_W_RESIDENT"]))
    if os.environ.get("PVSR_PACK_TABLE") is not None:     # A/B switch of the table-driven pack / scatter launches
        lib.pvsr_set_pack_table(int(os.environ["PVSR_PACK_TABLE"]))
    if os.environ.get("PVSR_TAIL_RANK1") is not None:     # A/B switch of the rank-1 adjoint of the head's tail
        lib.pvsr_set_tail_rank1(int(os.environ["PVSR_TAIL_RANK1"]))
    if os.environ.get("PVSR_TAIL_FWD") is not None:       # A/B switch of the composite forward of the head's tail
        lib.pvsr_set_tail_fwd(int(os.environ["PVSR_TAIL_FWD"]))
    if os.environ.get("PVSR_TWO_BRANCH") is not None:     # A/B switch of the two-branch training schedules
        lib.pvsr_set_two_branch(int(os.environ["PVSR_TWO_BRANCH"]))
    if os.environ.get("PVSR_PDL") is not None:            # A/B switch of programmatic dependent launch
        lib.pvsr_set_pdl(int(os.environ["PVSR_PDL"]))
    if os.environ.get("PVSR_CTA_PAIR") is not None:       # A/B switch of the cta_group::2 conv kernel
        lib.pvsr_set_cta_pair(int(os.environ["PVSR_CTA_PAIR"]))
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().pvsr_last_error()
        raise PvsrError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
