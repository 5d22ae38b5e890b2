"""ctypes binding of libtdvc_b200.so (the C ABI declared in include/tdvc_b200.h).

The library is built in-tree by build.sh / __graft_entry__.build().  There is no CPU or
PyTorch fallback: if the shared object is missing, or a tensor is not on a CUDA device, the
ops raise."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TDVC_LIB: another build of the same library (A/B measurements of compile-time variants); default: the in-tree build
LIB_PATH = os.environ.get("TDVC_LIB") or os.path.join(_HERE, "libtdvc_b200.so")

c_float_p = C.c_void_p  # device pointers travel as void*


class ConvGeom(C.Structure):
    """struct tdvc_conv_geom"""
    _fields_ = [("B", C.c_int32), ("Cin", C.c_int32), ("Tin", C.c_int32), ("Cout", C.c_int32),
                ("Tout", C.c_int32), ("K", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("dilation", C.c_int32), ("groups", C.c_int32), ("pad_mode", C.c_int32),
                ("in_slope", C.c_float), ("out_act", C.c_int32), ("out_slope", C.c_float)]


class TcConv(C.Structure):
    """struct tdvc_tc_conv"""
    _fields_ = [("xp", C.c_void_p), ("wp", C.c_void_p), ("bias", C.c_void_p), ("gb", C.c_void_p),
                ("residual", C.c_void_p), ("y", C.c_void_p), ("yp", C.c_void_p), ("maskp", C.c_void_p)] + \
               [(n, C.c_int32) for n in ("B", "Tp", "Tout", "K", "dilation", "t_off", "Cp_total", "groups", "a_ch_off",
                                         "a_ch_stride", "Cinp_g", "Cout_g", "Coutp_g", "bias_stride", "out_act")] + \
               [("out_slope", C.c_float)] + \
               [(n, C.c_int32) for n in ("out_packed", "tp_out", "cp_out", "out_halo", "out_ch_off", "out_ch_stride",
                                         "tm", "cm", "mask_halo", "mask_ch_off", "mask_ch_stride")] + \
               [("mask_slope", C.c_float), ("y_grp_stride", C.c_int64), ("y_b_stride", C.c_int64),
                ("chain_mode", C.c_int32), ("kg", C.c_int32 * 4), ("gb_grp_stride", C.c_int64), ("res_grp_stride", C.c_int64),
                ("yp2", C.c_void_p), ("auxp", C.c_void_p), ("dgbp", C.c_void_p), ("dgb_cp", C.c_int32),
                ("dgb_ch_off", C.c_int32), ("dgb_ch_stride", C.c_int32), ("halo_buf", C.c_void_p), ("halo", C.c_int32),
                ("t_valid", C.c_int32), ("pk_slope", C.c_float), ("unframe_s", C.c_int32), ("unframe_pad", C.c_int32),
                ("unframe_T", C.c_int32), ("unframe_C", C.c_int32), ("flat_tp", C.c_int32), ("flat_halo", C.c_int32),
                ("flat_T", C.c_int32)]


class PackJob(C.Structure):
    """struct tdvc_pack_job"""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p)] + \
               [(n, C.c_int32) for n in ("kind", "Cout", "Cin", "K", "Rp", "Qp", "flip", "R_total", "r_off", "Q_total", "q_off", "pad_")]


PACK_MAX_JOBS = 40


class L1Job(C.Structure):
    """struct tdvc_l1_job"""
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("da", C.c_void_p), ("n", C.c_int64), ("scale", C.c_float)]


L1_MAX_JOBS = 32


class TcWgrad2(C.Structure):
    """struct tdvc_tc_wgrad2"""
    _fields_ = [("dyp", C.c_void_p), ("xp", C.c_void_p), ("ws", C.c_void_p), ("dw", C.c_void_p * 4), ("db", C.c_void_p * 4),
                ("dw_grp_stride", C.c_int64), ("db_grp_stride", C.c_int64)] + \
               [(n, C.c_int32) for n in ("B", "Cdp", "Tout", "Cp", "Tp", "Cout", "Cin", "K", "dilation", "ngroups", "per_group",
                                         "x_ch_off", "x_ch_stride", "dy_ch_off", "dy_ch_stride")] + \
               [("kg", C.c_int32 * 4), ("t_off", C.c_int32 * 4)] + \
               [(n, C.c_int32) for n in ("want_bias", "ws_is_zero", "haloed", "tapsm", "swap", "frame_s", "kreal", "cin_conv_g", "sub")]


PAD_ZEROS, PAD_REFLECT = 0, 1
ACT_NONE, ACT_LRELU, ACT_TANH = 0, 1, 2

_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_G = C.POINTER(ConvGeom)

# name -> (restype, argtypes); mirrors include/tdvc_b200.h one to one (tests check every symbol)
SIGNATURES = {
    "tdvc_last_error": (C.c_char_p, []),
    "tdvc_version": (_I, []),
    "tdvc_launch_count": (_L, []),
    "tdvc_flop_count": (C.c_double, [_I]),
    "tdvc_device_is_sm100": (_I, []),
    "tdvc_weight_norm_fwd": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "tdvc_weight_norm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "tdvc_weight_norm_fwd_multi": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "tdvc_weight_norm_bwd_multi": (_I, [_P, _P, _I, _I, _P, _P, _P, _P]),
    "tdvc_conv1d_fwd": (_I, [_G, _P, _P, _P, _P, _P, _P]),
    "tdvc_conv1d_bwd_data_ws": (_L, [_G]),
    "tdvc_conv1d_bwd_data": (_I, [_G, _P, _P, _P, _P, _P, _P]),
    "tdvc_bias_grad": (_I, [_P, _P, _I, _I, _I, _P]),
    "tdvc_pad_act_bwd": (_I, [_P, _P, _P, _L, _I, _I, _I, _F, _P]),
    "tdvc_conv1d_bwd_weight": (_I, [_G, _P, _P, _P, _P, _P]),
    "tdvc_conv_transpose1d_fwd": (_I, [_G, _P, _P, _P, _P, _P]),
    "tdvc_conv_transpose1d_bwd_data": (_I, [_G, _P, _P, _P, _P]),
    "tdvc_conv_transpose1d_bwd_weight": (_I, [_G, _P, _P, _P, _P, _P]),
    "tdvc_leaky_relu_fwd": (_I, [_P, _P, _L, _F, _P]),
    "tdvc_f0_unvoiced_count": (_I, [_P, _P, _I, _I, _I, _F, _I, _P]),
    "tdvc_f0_excitation": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _P]),
    "tdvc_yin_estimate": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
    "tdvc_gate_fwd": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "tdvc_gate_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "tdvc_act_bwd_from_output": (_I, [_P, _P, _P, _L, _I, _F, _P]),
    "tdvc_film_fwd": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "tdvc_film_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "tdvc_add3_scale": (_I, [_P, _P, _P, _P, _L, _F, _P]),
    "tdvc_l2norm_fwd": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "tdvc_l2norm_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "tdvc_cond_concat_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "tdvc_cond_concat_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "tdvc_instnorm_stats": (_I, [_P, _P, _P, _I, _I, _F, _P]),
    "tdvc_cin_apply_fwd": (_I, [_P, _P, _P, _P, _I, _P, _I, _I, _I, _F, _P]),
    "tdvc_cin_apply_bwd": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _F, _P]),
    "tdvc_avgpool4s2_fwd": (_I, [_P, _P, _I, _I, _I, _P]),
    "tdvc_avgpool4s2_bwd": (_I, [_P, _P, _I, _I, _I, _P]),
    "tdvc_select_channel_fwd": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "tdvc_select_channel_bwd": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "tdvc_conv1d_select_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "tdvc_conv1d_select_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "tdvc_sq_err_const_sum": (_I, [_P, _F, _F, _P, _L, _P]),
    "tdvc_sq_err_const_bwd": (_I, [_P, _F, _F, _P, _P, _L, _P]),
    "tdvc_abs_diff_sum": (_I, [_P, _P, _F, _P, _L, _P]),
    "tdvc_abs_diff_bwd": (_I, [_P, _P, _F, _P, _P, _L, _P]),
    "tdvc_abs_diff_sum_multi": (_I, [C.POINTER(L1Job), _I, _P, _P]),
    "tdvc_abs_diff_bwd_multi": (_I, [C.POINTER(L1Job), _I, _P, _P]),
    "tdvc_contrastive_dir": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "tdvc_adamw_multi": (_I, [_P, _P, _P, _P, _P, _I, _L, _F, _F, _F, _F, _F, _I, _F, _P, _P]),
    "tdvc_adamw_blocks": (_I, [_P, _P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _F, _I, _F, _P, _P]),
    "tdvc_pack_cl_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _I, _I, _I, _P, _P]),
    "tdvc_cond_pack_cl": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "tdvc_pack_jobs": (_I, [_P, _I, _P]),
    "tdvc_pack_cl_bf16_masked": (_I, [_P, _P, _F, _P, _I, _I, _I, _I, _I, _P, _P]),
    "tdvc_pack_weight_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "tdvc_conv1d_tc_wgrad_ws": (_L, [_I, _I, _I]),
    "tdvc_conv1d_tc_wgrad": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    "tdvc_pack_weight_bf16_multi": (_I, [_P, _I, _I, _P, _P, _P]),
    "tdvc_space_to_depth": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "tdvc_depth_to_space": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "tdvc_conv1d_tc_fwd_ex": (_I, [C.POINTER(TcConv), _P]),
    "tdvc_chain_fold": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _L, _L, _I, _I, _I, _P]),
    "tdvc_conv1d_tc_wgrad2_ws": (_L, [C.POINTER(TcWgrad2)]),
    "tdvc_conv1d_tc_wgrad2": (_I, [C.POINTER(TcWgrad2), _P]),
    "tdvc_stft_frames_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "tdvc_stft_frames_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "tdvc_power_fwd": (_I, [_P, _P, _I, _I, _L, _I, _P]),
    "tdvc_power_bwd": (_I, [_P, _P, _P, _I, _I, _I, _L, _I, _P]),
    "tdvc_log_clamp_fwd": (_I, [_P, _P, _L, _F, _P]),
    "tdvc_log_clamp_bwd": (_I, [_P, _P, _P, _L, _F, _P]),
    "tdvc_chain_pack": (_I, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int), _I, _I, _I, _P, _P, _P, _P]),
    "tdvc_frame_weights_pack": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "tdvc_frame_pack_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "tdvc_frame_unpack": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "tdvc_conv1d_tc_fwd_stacked": (_I, [_P, _P, _P, _P] + [_I] * 12 + [_I, _F] + [_I] * 5 + [_P]),
    "tdvc_conv1d_tc_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P]),
}

_lib = None


def load():
    """Loads (once) and returns the CDLL.  Raises RuntimeError when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"tdvc: {LIB_PATH} is missing -- build it with ./build.sh (or __graft_entry__.build()); "
            "there is no CPU / PyTorch fallback for the hot path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().tdvc_last_error()
        raise RuntimeError(f"tdvc {what} failed ({rc}): {msg.decode() if msg else '?'}")
