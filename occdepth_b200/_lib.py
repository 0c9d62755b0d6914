"""ctypes binding of include/occdepth_b200.h (the C-ABI drop-in boundary).

The library is mandatory: there is no CPU or PyTorch fallback behind these calls.  `lib()` raises if
libocc_b200.so has not been built (python -m occdepth_b200._build / __graft_entry__.build()).
"""
import ctypes as C
import os

from . import _build

_LIB = None

MAX_SCALES = 4
DTYPE_F32, DTYPE_BF16 = 0, 1
SFA_OUT_F32_PLANAR, SFA_OUT_BF16_CL, SFA_OUT_F32_CL, SFA_OUT_TF32_CL = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SILU, ACT_SIGMOID = 0, 1, 2, 3, 4


class SfaParams(C.Structure):
    _fields_ = [
        ("feat", C.c_void_p * MAX_SCALES),
        ("h", C.c_int * MAX_SCALES),
        ("w", C.c_int * MAX_SCALES),
        ("div", C.c_int * MAX_SCALES),
        ("vstride", C.c_longlong * MAX_SCALES),
        ("n_scales", C.c_int),
        ("n_views", C.c_int),
        ("C", C.c_int),
        ("feat_dtype", C.c_int),
        ("pix", C.c_void_p),
        ("fov", C.c_void_p),
        ("N", C.c_longlong),
        ("P", C.c_int),
        ("out", C.c_void_p),
        ("out_mode", C.c_int),
        ("out_cstride", C.c_int),
        ("perm_nyu", C.c_int),
        ("S1", C.c_int),
        ("S2", C.c_int),
        ("prior", C.c_void_p),
        ("scale_const", C.c_float),
    ]


CONV_MAX_TAPS, CONV_MAX_SRC, CONV_MAX_GROUPS = 81, 3, 8
CONV_IMPL_TC, CONV_IMPL_SIMT, CONV_IMPL_HALO, CONV_IMPL_HALOX, CONV_IMPL_TCX = 0, 1, 2, 3, 4
OUT1_NONE, OUT1_CL, OUT1_F32_PLANAR = 0, 1, 2


class ConvTap(C.Structure):
    _fields_ = [("src", C.c_int), ("dz", C.c_int), ("dy", C.c_int), ("dx", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("impl", C.c_int),
        ("dtype", C.c_int),
        ("n_src", C.c_int),
        ("src", C.c_void_p * CONV_MAX_SRC),
        ("src_C", C.c_int * CONV_MAX_SRC),
        ("src_cstride", C.c_int * CONV_MAX_SRC),
        ("src_coff", C.c_int * CONV_MAX_SRC),
        ("B", C.c_int), ("ID", C.c_int), ("IH", C.c_int), ("IW", C.c_int),
        ("src_d0", C.c_int),
        ("stride", C.c_int * 3),
        ("n_taps", C.c_int),
        ("taps", ConvTap * CONV_MAX_TAPS),
        ("weight", C.c_void_p),
        ("bias", C.c_void_p),
        ("Cout", C.c_int), ("Cout_pad", C.c_int), ("Kpad", C.c_int),
        ("weight_per_image", C.c_int),
        ("OD", C.c_int), ("OH", C.c_int), ("OW", C.c_int),
        ("omul", C.c_int * 3), ("oadd", C.c_int * 3),
        ("ODf", C.c_int), ("OHf", C.c_int), ("OWf", C.c_int),
        ("out0", C.c_void_p),
        ("out0_cstride", C.c_int), ("out0_coff", C.c_int),
        ("act", C.c_int),
        ("out0_exact", C.c_int),
        ("res1", C.c_void_p),
        ("res1_cstride", C.c_int), ("res1_coff", C.c_int),
        ("res2", C.c_void_p),
        ("res2_cstride", C.c_int), ("res2_coff", C.c_int),
        ("res2_post", C.c_int),
        ("out1_mode", C.c_int),
        ("out1", C.c_void_p),
        ("out1_cstride", C.c_int), ("out1_coff", C.c_int),
        ("out1_C", C.c_int),
        ("n_groups", C.c_int),
        ("group_tap0", C.c_int * (CONV_MAX_GROUPS + 1)),
        ("group_oadd", (C.c_int * 3) * CONV_MAX_GROUPS),
    ]


# every symbol include/occdepth_b200.h declares: name -> (restype, argtypes)
_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
SYMBOLS = {
    "occd_abi_version": (C.c_int, []),
    "occd_last_error": (C.c_char_p, []),
    "occd_sfa_lift_fwd": (C.c_int, [C.POINTER(SfaParams), _vp]),
    "occd_planar_to_cl": (C.c_int, [_vp, _vp, _i, _ll, _i, _ll, _i, _vp]),
    "occd_cl_to_planar": (C.c_int, [_vp, _i, _vp, _ll, _i, _ll, _i, _vp]),
    "occd_conv_plan_create": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_void_p)]),
    "occd_conv_plan_destroy": (C.c_int, [_vp]),
    "occd_conv_run": (C.c_int, [_vp, _vp]),
    "occd_conv_plan_info": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "occd_conv_debug_trace": (C.c_int, [_vp]),
    "occd_softmax_planar_to_cl": (C.c_int, [_vp, _vp, _i, _ll, _i, _ll, _i, _i, _vp]),
    "occd_argmax_classes": (C.c_int, [_vp, _vp, _ll, _i, _ll, _vp, _vp]),
    "occd_normalize_rgb_u8": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "occd_vox2pix_fwd": (C.c_int, [_vp, _i, _vp, _vp, C.c_double, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "occd_cl_transpose": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _ll, _vp]),
    "occd_copy_channels": (C.c_int, [_vp, _vp, _i, _ll, _i, _i, _i, _i, _i, _vp]),
    "occd_dwconv2d_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp] + [_i] * 14 + [_vp]),
    "occd_dwconv2d_tiled_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp] + [_i] * 14 + [_vp]),
    "occd_se_gate_fwd": (C.c_int, [_vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "occd_se_gate_fold_fwd": (C.c_int, [_vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "occd_scale_weights": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "occd_frustum_sample_fwd": (C.c_int, [_vp, _vp] + [_i] * 7 + [_f] * 4 + [_i, _vp, _i, _vp]),
    "occd_grid_sample_prior_fwd": (C.c_int, [_vp, _vp] + [_i] * 8 + [_vp, _i, _vp]),
    "occd_softmax_planar": (C.c_int, [_vp, _vp, _ll, _i, _ll, _vp]),
    "occd_fc_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "occd_channel_scale": (C.c_int, [_vp, _vp, _i, _ll, _ll, _i, _i, _vp]),
    "occd_virtual_view_fwd": (C.c_int, [_vp, _vp, _vp] + [_i] * 9 + [_f, _vp]),
    "occd_upsample_bilinear_ac": (C.c_int, [_vp, _vp] + [_i] * 11 + [_vp]),
}


def lib():
    """Load libocc_b200.so (once).  Fails loudly when it is missing -- never falls back."""
    global _LIB
    if _LIB is None:
        path = _build.lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                "occdepth_b200: %s is missing -- build it with `python -m occdepth_b200._build` "
                "(or __graft_entry__.build()); there is no CPU fallback" % path)
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = lib().occd_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, (msg or b"").decode()))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
