"""ctypes loader of the C-ABI library (include/vmambair_b200.h).

There is NO fallback: if the CUDA library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VMB_LIB_PATH") or os.path.join(_HERE, "lib", "libvmambair_b200.so")
_lib = None

DT_F32, DT_BF16, DT_F16 = 0, 1, 2

i64 = C.c_int64
vp = C.c_void_p


class ScanFwdArgs(C.Structure):
    _fields_ = [
        ("u", vp), ("delta", vp), ("A", vp), ("Bm", vp), ("Cm", vp), ("D", vp), ("delta_bias", vp),
        ("out", vp), ("ckpt", vp),
        ("batch", C.c_int), ("dim", C.c_int), ("seqlen", C.c_int), ("dstate", C.c_int), ("ngroups", C.c_int),
        ("u_bs", i64), ("u_ds", i64), ("delta_bs", i64), ("delta_ds", i64), ("out_bs", i64), ("out_ds", i64),
        ("B_bs", i64), ("B_gs", i64), ("B_ns", i64), ("C_bs", i64), ("C_gs", i64), ("C_ns", i64),
        ("delta_softplus", C.c_int), ("dtype", C.c_int),
    ]


class ScanBwdArgs(C.Structure):
    _fields_ = [
        ("u", vp), ("delta", vp), ("A", vp), ("Bm", vp), ("Cm", vp), ("D", vp), ("delta_bias", vp),
        ("dout", vp), ("ckpt", vp),
        ("du", vp), ("ddelta", vp), ("dA", vp), ("dB", vp), ("dC", vp), ("dD", vp), ("ddelta_bias", vp), ("workspace", vp),
        ("batch", C.c_int), ("dim", C.c_int), ("seqlen", C.c_int), ("dstate", C.c_int), ("ngroups", C.c_int),
        ("u_bs", i64), ("u_ds", i64), ("delta_bs", i64), ("delta_ds", i64), ("dout_bs", i64), ("dout_ds", i64),
        ("du_bs", i64), ("du_ds", i64), ("ddelta_bs", i64), ("ddelta_ds", i64),
        ("B_bs", i64), ("B_gs", i64), ("B_ns", i64), ("C_bs", i64), ("C_gs", i64), ("C_ns", i64),
        ("delta_softplus", C.c_int), ("dtype", C.c_int),
    ]


class PixlinArgs(C.Structure):
    _fields_ = [
        ("x", vp), ("w", vp), ("bias", vp), ("residual", vp), ("out", vp), ("ln_w", vp), ("ln_b", vp), ("gate", vp),
        ("ln_mode", C.c_int), ("gate_mode", C.c_int), ("act_from", C.c_int), ("act_to", C.c_int),
        ("batch", C.c_int), ("K", C.c_int), ("M", C.c_int), ("P", C.c_int),
        ("x_bs", i64), ("x_cs", i64), ("r_bs", i64), ("r_cs", i64), ("o_bs", i64), ("o_cs", i64), ("g_bs", i64), ("w_ld", i64),
        ("dtype", C.c_int), ("out_dtype", C.c_int), ("w_static", C.c_int),
    ]


class DwconvArgs(C.Structure):
    _fields_ = [
        ("x", vp), ("w", vp), ("bias", vp), ("out", vp),
        ("batch", C.c_int), ("c_out", C.c_int), ("H", C.c_int), ("W", C.c_int), ("mode", C.c_int),
        ("x_bs", i64), ("x_cs", i64), ("o_bs", i64), ("o_cs", i64), ("dtype", C.c_int),
    ]


class CrossScanArgs(C.Structure):
    _fields_ = [
        ("src", vp * 4), ("out", vp),
        ("batch", C.c_int), ("rows", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("src_bs", i64), ("src_rs", i64), ("out_bs", i64), ("dtype", C.c_int), ("out_ks", i64),
    ]


class MergeArgs(C.Structure):
    _fields_ = [
        ("ys", vp), ("z", vp), ("ln_w", vp), ("ln_b", vp), ("y2", vp), ("pooled", vp),
        ("batch", C.c_int), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("z_bs", i64), ("z_cs", i64), ("dtype", C.c_int), ("workspace", vp), ("in_place_order", C.c_int),
        ("z_preact", C.c_int), ("save_ws", C.c_int),
    ]


class TransposeArgs(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("planes", C.c_int), ("H", C.c_int), ("W", C.c_int), ("dtype", C.c_int)]


class PixelShuffleArgs(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("batch", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("dtype", C.c_int)]


class Conv3x3Args(C.Structure):
    _fields_ = [
        ("x", vp), ("w", vp), ("bias", vp), ("out", vp), ("add", vp),
        ("batch", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("in_nhwc", C.c_int), ("mode", C.c_int), ("add_scale", C.c_int),
        ("x_bs", i64), ("x_cs", i64), ("o_bs", i64), ("o_cs", i64), ("add_bs", i64), ("add_cs", i64),
        ("dtype", C.c_int),
    ]


class ScanGroupedArgs(C.Structure):
    _fields_ = [
        ("u", vp * 4), ("delta", vp * 4), ("Bm", vp * 4), ("Cm", vp * 4), ("out", vp * 4), ("rev", C.c_int * 4),
        ("A", vp), ("D", vp), ("delta_bias", vp),
        ("batch", C.c_int), ("dim", C.c_int), ("seqlen", C.c_int), ("dstate", C.c_int), ("ngroups", C.c_int),
        ("u_bs", i64), ("u_ds", i64), ("delta_bs", i64), ("delta_ds", i64), ("out_bs", i64), ("out_ds", i64),
        ("B_bs", i64), ("B_ns", i64), ("C_bs", i64), ("C_ns", i64),
        ("delta_softplus", C.c_int), ("dtype", C.c_int),
    ]


class ChannelArgs(C.Structure):
    _fields_ = [
        ("pooled", vp), ("inv_count", C.c_float),
        ("cin_w", vp), ("cin_b", vp), ("xc_proj", vp), ("dtc_w", vp), ("dtc_b", vp), ("Ac_logs", vp), ("Dsc", vp),
        ("cout_w", vp), ("cout_b", vp), ("cn_w", vp), ("cn_b", vp), ("c_out", vp),
        ("batch", C.c_int), ("C", C.c_int), ("dc", C.c_int), ("Rc", C.c_int), ("N", C.c_int),
    ]


class LnFwdArgs(C.Structure):
    _fields_ = [("x", vp), ("w", vp), ("b", vp), ("y", vp), ("stats", vp),
                ("batch", C.c_int), ("C", C.c_int), ("L", C.c_int), ("mode", C.c_int),
                ("x_bs", i64), ("x_cs", i64), ("y_bs", i64), ("y_cs", i64), ("dtype", C.c_int)]


class LnBwdArgs(C.Structure):
    _fields_ = [("x", vp), ("g", vp), ("add", vp), ("w", vp), ("dx", vp), ("dw", vp), ("db", vp), ("stats", vp),
                ("batch", C.c_int), ("C", C.c_int), ("L", C.c_int), ("mode", C.c_int),
                ("x_bs", i64), ("x_cs", i64), ("g_bs", i64), ("g_cs", i64), ("a_bs", i64), ("a_cs", i64),
                ("dx_bs", i64), ("dx_cs", i64), ("dtype", C.c_int)]


class MergeBwdArgs(C.Structure):
    _fields_ = [("merged", vp), ("stats", vp), ("z", vp), ("dy2", vp), ("dpooled", vp), ("w", vp), ("b", vp),
                ("dm", vp), ("dz", vp), ("dw", vp), ("db", vp),
                ("batch", C.c_int), ("C", C.c_int), ("L", C.c_int),
                ("z_bs", i64), ("z_cs", i64), ("dz_bs", i64), ("dz_cs", i64), ("dtype", C.c_int)]


class DwconvBwdArgs(C.Structure):
    _fields_ = [("x", vp), ("w", vp), ("bias", vp), ("g", vp), ("dv", vp), ("dw", vp), ("dbias", vp),
                ("batch", C.c_int), ("c_out", C.c_int), ("H", C.c_int), ("W", C.c_int), ("mode", C.c_int),
                ("x_bs", i64), ("x_cs", i64), ("g_bs", i64), ("g_cs", i64), ("dv_bs", i64), ("dv_cs", i64), ("dtype", C.c_int)]


class GateBwdArgs(C.Structure):
    _fields_ = [("dyg", vp), ("y2", vp), ("gate", vp), ("dy2", vp), ("dgate", vp),
                ("batch", C.c_int), ("C", C.c_int), ("L", C.c_int), ("mode", C.c_int), ("dtype", C.c_int)]


class WgradArgs(C.Structure):
    _fields_ = [("dy", vp), ("x", vp), ("out", vp), ("batch", C.c_int), ("M", C.c_int), ("K", C.c_int), ("L", C.c_int),
                ("dy_bs", i64), ("dy_cs", i64), ("x_bs", i64), ("x_cs", i64), ("per_batch", C.c_int), ("dtype", C.c_int),
                ("dbias", vp)]


class ChannelBwdArgs(C.Structure):
    _fields_ = [("fwd", ChannelArgs), ("dc_out", vp), ("d_pooled", vp),
                ("d_cin_w", vp), ("d_cin_b", vp), ("d_xc_proj", vp), ("d_dtc_w", vp), ("d_dtc_b", vp), ("d_Ac_logs", vp), ("d_Dsc", vp),
                ("d_cout_w", vp), ("d_cout_b", vp), ("d_cn_w", vp), ("d_cn_b", vp), ("scratch", vp)]


class PrepJob(C.Structure):
    _fields_ = [("src", vp), ("src2", vp), ("dst", vp), ("dst2", vp), ("type", C.c_int), ("M", C.c_int), ("K", C.c_int),
                ("N2", C.c_int), ("ld", C.c_int), ("ld2", C.c_int)]


class PrepArgs(C.Structure):
    _fields_ = [("jobs", PrepJob * 16), ("njobs", C.c_int), ("dtype", C.c_int)]


class AdamArgs(C.Structure):
    _fields_ = [("param", vp), ("grad", vp), ("exp_avg", vp), ("exp_avg_sq", vp), ("ema", vp), ("state", vp),
                ("n", C.c_long), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("decoupled_weight_decay", C.c_int),
                ("grad_scale", C.c_float), ("max_grad_norm", C.c_float), ("ema_decay", C.c_float), ("zero_grad", C.c_int)]


# every symbol include/vmambair_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "vmb_last_error": (C.c_char_p, []),
    "vmb_version": (C.c_char_p, []),
    "vmb_scan_ckpt_interval": (C.c_int, []),
    "vmb_selective_scan_fwd": (C.c_int, [C.POINTER(ScanFwdArgs), vp]),
    "vmb_selective_scan_bwd": (C.c_int, [C.POINTER(ScanBwdArgs), vp]),
    "vmb_scan_bwd_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "vmb_pixlin": (C.c_int, [C.POINTER(PixlinArgs), vp]),
    "vmb_dwconv3x3": (C.c_int, [C.POINTER(DwconvArgs), vp]),
    "vmb_cross_scan": (C.c_int, [C.POINTER(CrossScanArgs), vp]),
    "vmb_merge_norm_gate": (C.c_int, [C.POINTER(MergeArgs), vp]),
    "vmb_merge_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "vmb_transpose_hw": (C.c_int, [C.POINTER(TransposeArgs), vp]),
    "vmb_pixel_shuffle2_nhwc": (C.c_int, [C.POINTER(PixelShuffleArgs), vp]),
    "vmb_pixel_shuffle2_nhwc_bias": (C.c_int, [C.POINTER(PixelShuffleArgs), vp, vp]),
    "vmb_conv3x3": (C.c_int, [C.POINTER(Conv3x3Args), vp]),
    "vmb_cross_scan_multi": (C.c_int, [C.POINTER(CrossScanArgs), C.c_int, vp]),
    "vmb_dwconv3x3_t": (C.c_int, [C.POINTER(DwconvArgs), vp, vp]),
    "vmb_selective_scan_fwd_grouped": (C.c_int, [C.POINTER(ScanGroupedArgs), vp]),
    "vmb_channel_branch": (C.c_int, [C.POINTER(ChannelArgs), vp]),
    "vmb_layernorm_fwd": (C.c_int, [C.POINTER(LnFwdArgs), vp]),
    "vmb_layernorm_bwd": (C.c_int, [C.POINTER(LnBwdArgs), vp]),
    "vmb_merge_norm_gate_bwd": (C.c_int, [C.POINTER(MergeBwdArgs), vp]),
    "vmb_dwconv3x3_bwd": (C.c_int, [C.POINTER(DwconvBwdArgs), vp]),
    "vmb_channel_gate_bwd": (C.c_int, [C.POINTER(GateBwdArgs), vp]),
    "vmb_fused_adam": (C.c_int, [C.POINTER(AdamArgs), vp]),
    "vmb_pixlin_wgrad": (C.c_int, [C.POINTER(WgradArgs), vp]),
    "vmb_channel_branch_bwd": (C.c_int, [C.POINTER(ChannelBwdArgs), vp]),
    "vmb_channel_branch_bwd_smem_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "vmb_channel_branch_bwd_scratch_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "vmb_prep_block_weights": (C.c_int, [C.POINTER(PrepArgs), vp]),
    "vmb_sum4_add": (C.c_int, [vp, vp, vp, C.c_int, C.c_long, C.c_int, vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"vmambair_b200: CUDA library not built ({LIB_PATH}); run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C vmambair_b200/csrc`. There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int, what: str):
    if status != 0:
        msg = lib().vmb_last_error().decode()
        raise RuntimeError(f"{what}: {msg} (status {status})")
