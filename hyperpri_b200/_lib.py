"""ctypes binding of the C ABI declared in include/hyperpri_b200.h.

The shared library is built in-tree by ``hyperpri_b200.build`` (nvcc, sm_100a) and MUST be
present: there is no CPU or PyTorch fallback -- a missing library raises at first use.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhyperpri_b200.so")


class Conv3x3Job(C.Structure):
    """hpri_conv3x3_job_t (include/hyperpri_b200.h)."""
    _fields_ = [("w", C.c_void_p), ("dst_fwd", C.c_void_p), ("dst_dgrad", C.c_void_p), ("grad_packed", C.c_void_p),
                ("grad_dst", C.c_void_p), ("cout", C.c_int), ("cin", C.c_int), ("fwd_dtype", C.c_int),
                ("dgrad_dtype", C.c_int), ("tile0", C.c_int), ("kind", C.c_int)]


class AdamJob(C.Structure):
    """hpri_adam_job_t (include/hyperpri_b200.h)."""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_longlong), ("block0", C.c_int), ("pad_", C.c_int)]


class BnFin(C.Structure):
    """hpri_bn_fin_t (include/hyperpri_b200.h)."""
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p), ("conv_bias", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p), ("scale", C.c_void_p),
                ("shift", C.c_void_p), ("save_mean", C.c_void_p), ("save_invstd", C.c_void_p),
                ("counter", C.c_void_p), ("count", C.c_longlong), ("momentum", C.c_float), ("eps", C.c_float),
                ("partials", C.c_void_p)]


class View(C.Structure):
    """hpri_view_t: NHWC bf16 view with element strides."""
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("c", C.c_int),
                ("pix_stride", C.c_longlong), ("row_stride", C.c_longlong), ("img_stride", C.c_longlong),
                ("dtype", C.c_int)]


_VP = C.POINTER(View)


class BnBwd(C.Structure):
    """hpri_bn_bwd_t (include/hyperpri_b200.h)."""
    _fields_ = [("x", _VP), ("scale", C.c_void_p), ("shift", C.c_void_p), ("save_mean", C.c_void_p),
                ("save_invstd", C.c_void_p), ("sums", C.c_void_p)]

_p, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes; every symbol include/hyperpri_b200.h declares
SIGNATURES = {
    "hpri_abi_version": [],
    "hpri_launch_count": [],
    "hpri_igemm_fwd": [_VP, _p, _i, _i, _i, _i, _VP, _i, _p, _p, _i, _i, C.POINTER(BnFin), C.POINTER(BnBwd), _p],
    "hpri_conv3x3_halo_ok": [_i, _i, _i],
    "hpri_set_conv_algo": [_i],
    "hpri_set_wgrad_algo": [_i],
    "hpri_set_sm_reserve": [_i],
    "hpri_set_halo_a_stages": [_i],
    "hpri_set_reverse_elementwise": [_i],
    "hpri_set_deterministic": [_i],
    "hpri_convT2x2_fwd": [_VP, _p, _i, _i, _i, _VP, _p, _i, _p],
    "hpri_convT2x2_dgrad": [_VP, _p, _i, _i, _i, _VP, _i, _p],
    "hpri_igemm_wgrad": [_VP, _VP, _i, _i, _p, _i, _i, _i, _p],
    "hpri_pack_weights": [_p, _p, _i, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _p],
    "hpri_unpack_grads": [_p, _p, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _f, _i, _f, _p, _p],
    "hpri_pack_conv3x3": [_p, _i, _i, _p, _i, _p, _i, _p],
    "hpri_unpack_conv3x3": [_p, _i, _i, _p, _p],
    "hpri_pr_hist": [_p, _p, _ll, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p],
    "hpri_adam_step": [_p, _i, _i, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _i, _p, _p],
    "hpri_pack_conv3x3_batch": [_p, _i, _i, _p],
    "hpri_unpack_conv3x3_batch": [_p, _i, _i, _f, _p, _p],
    "hpri_pack_convT2x2": [_p, _i, _i, _p, _i, _p],
    "hpri_unpack_convT2x2": [_p, _i, _i, _p, _p],
    "hpri_hsi_ingest": [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _i, _i, _p],
    "hpri_hsi_ingest_f16": [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _i, _i, _p],
    "hpri_absmax": [_p, _ll, _p, _p],
    "hpri_convert16": [_VP, _VP, _p],
    "hpri_mul16": [_VP, _VP, _VP, _p],
    "hpri_upsample2_fwd": [_VP, _VP, _p],
    "hpri_upsample2_bwd": [_VP, _VP, _p],
    "hpri_bn_finalize": [_p, _ll, _p, _p, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p, _p, _i, _p],
    "hpri_bn_relu_apply": [_VP, _p, _p, _VP, _VP, _p],
    "hpri_bn_relu_bwd_reduce": [_VP, _p, _p, _p, _p, _VP, _VP, _p, _p, _p, _p],
    "hpri_bn_relu_bwd_apply": [_VP, _p, _p, _p, _p, _p, _VP, _VP, _p, _p, _p, _ll, _VP, _p, _p, _p, _f, _f, _p, _p],
    "hpri_head_fwd": [_VP, _p, _p, _p, _p, _p, _p],
    "hpri_bce_fwd_bwd": [_p, _p, _ll, _f, _f, _p, _p, _p, _p],
    "hpri_colsum": [_VP, _p, _f, _f, _p],
    "hpri_sum_f32": [_p, _ll, _p, _f, _p],
    "hpri_scale_check": [_p, _ll, _f, _p, _p],
}

ERRORS = {-1: "HPRI_ERR_ARG", -2: "HPRI_ERR_ALIGN", -3: "HPRI_ERR_DRIVER", -4: "HPRI_ERR_TENSORMAP",
          -5: "HPRI_ERR_CUDA"}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load (once) and return the CDLL; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: run `python -m hyperpri_b200.build` (nvcc, sm_100a). "
                "hyperpri_b200 has no CPU/PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the .so is stale
            fn.argtypes = args
            fn.restype = C.c_longlong if name == "hpri_launch_count" else C.c_int
        _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {ERRORS.get(rc, rc)}")
