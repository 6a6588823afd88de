"""ctypes binding of libsynthanatomy_b200.so (the C ABI declared in include/synthanatomy_b200.h).

There is NO fallback: if the shared library is missing and cannot be built, importing raises; if a call
fails, a RuntimeError carrying ``sa_last_error()`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

c_void_p = C.c_void_p
c_int = C.c_int
c_int64 = C.c_int64
c_float = C.c_float
c_double = C.c_double

SA_F32, SA_BF16 = 0, 1
SA_PATH_NONE, SA_PATH_SIMT, SA_PATH_TCGEN05 = 0, 1, 2


class ConvDesc(C.Structure):
    """mirror of `sa_conv_desc`"""
    _fields_ = [
        ("batch", C.c_int32),
        ("in_dhw", C.c_int32 * 3),
        ("out_dhw", C.c_int32 * 3),
        ("c_in", C.c_int32),
        ("c_out", C.c_int32),
        ("ksize", C.c_int32),
        ("stride", C.c_int32),
        ("pad", C.c_int32),
        ("transposed", C.c_int32),
        ("act_dtype", C.c_int32),
    ]


class GemmEpilogue(C.Structure):
    """mirror of `sa_gemm_epilogue` (include/synthanatomy_b200_performer.h)"""
    _fields_ = [
        ("bias", c_void_p),
        ("dot_with", c_void_p),
        ("dot_out", c_void_p),
        ("scale_dev", c_void_p),
        ("scale", C.c_float),
        ("act", C.c_int32),
        ("pre", c_void_p),
        ("resid", c_void_p),
        ("out_f32", c_void_p),
        ("out_act", c_void_p),
    ]


class FavorDesc(C.Structure):
    """mirror of `sa_favor_desc`"""
    _fields_ = [(n, C.c_int32) for n in ("batch", "seq", "heads", "dim_head", "m", "mp", "ld", "act_dtype")]


class LocalDesc(C.Structure):
    """mirror of `sa_local_desc`"""
    _fields_ = [(n, C.c_int32) for n in ("batch", "seq", "heads", "dim_head", "window", "ld", "out_ld", "act_dtype")]


SA_ACT_NONE, SA_ACT_GELU_FWD, SA_ACT_GELU_BWD, SA_ACT_GELU_FWD_D, SA_ACT_MUL_PRE = 0, 1, 2, 3, 4

# name -> (restype, argtypes); the single source of truth for tests/test_abi.py as well
class WPackItem(C.Structure):
    """sa_wpack_item (include/synthanatomy_b200.h)"""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("A", C.c_int), ("B", C.c_int), ("taps", C.c_int),
                ("transpose", C.c_int), ("flip", C.c_int)]


class WPrepItem(C.Structure):
    """sa_wprep_item (include/synthanatomy_b200.h)"""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("dst_t", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int),
                ("dst_ld", C.c_int), ("dst_t_ld", C.c_int)]


SIGNATURES = {
    "sa_last_error": (C.c_char_p, []),
    "sa_version": (c_int, []),
    "sa_last_path": (c_int, []),
    "sa_launch_count": (c_int64, []),
    "sa_launch_count_reset": (None, []),
    "sa_set_force_simt": (None, [c_int]),
    "sa_set_deterministic": (None, [c_int]),
    "sa_get_deterministic": (c_int, []),
    "sa_conv3d_fwd": (c_int, [C.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                              c_void_p]),
    "sa_conv3d_wgrad": (c_int, [C.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "sa_pack_weight": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "sa_unpack_wgrad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "sa_im2col_c1": (c_int, [c_void_p, c_int, c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), c_int, c_int, c_int, c_void_p,
                             c_void_p]),
    "sa_col2im_c1": (c_int, [c_void_p, c_int, c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), c_int, c_int, c_int, c_void_p,
                             c_void_p, c_void_p]),
    "sa_bias_grad": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p]),
    "sa_vq_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                              c_void_p, c_void_p]),
    "sa_vq_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int64, c_void_p, c_void_p]),
    "sa_vq_ema_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_double,
                                 c_void_p, c_void_p]),
    "sa_vq_perplexity": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "sa_vq_embed": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "sa_nchw_to_nhwc": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_int64, c_void_p]),
    "sa_nhwc_to_nchw": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_int64, c_void_p]),
    "sa_cast": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p]),
    "sa_weight_prep": (c_int, [c_void_p, c_int, c_void_p]),
    "sa_pack_weight_multi": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "sa_bn_workspace": (C.c_size_t, [c_int]),
    "sa_bn_stats": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p]),
    "sa_bn_eval_stats": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "sa_bn_lrelu_fwd": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                c_void_p]),
    "sa_bn_lrelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sa_lrelu_fwd": (c_int, [c_void_p, c_int, c_int64, c_float, c_void_p]),
    "sa_lrelu_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_float, c_void_p]),
    "sa_mse_fwd_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sa_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_int,
                             c_void_p]),
    "sa_swap_outer_inner": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "sa_spectral_amp_loss": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sa_adam_multi": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float,
                              c_int, c_void_p]),
    "sa_conv3d_x3_supported": (c_int, [C.POINTER(ConvDesc), c_int]),
    "sa_conv3d_x3_workspace": (C.c_size_t, [C.POINTER(ConvDesc), c_int]),
    "sa_conv3d_fwd_x3": (c_int, [C.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                 c_void_p, C.c_size_t, c_void_p]),
    "sa_conv3d_wgrad_x3": (c_int, [C.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_int, c_void_p, C.c_size_t, c_void_p]),
    # ---- Performer path (include/synthanatomy_b200_performer.h)
    "sa_gemm_nt_x3_workspace": (C.c_size_t, [c_int64, c_int, c_int]),
    "sa_gemm_nt_x3": (c_int, [c_int64, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, C.POINTER(GemmEpilogue), c_int64,
                              c_void_p, C.c_size_t, c_void_p]),
    "sa_gemm_tn_x3_workspace": (C.c_size_t, [c_int64, c_int, c_int]),
    "sa_gemm_tn_x3": (c_int, [c_int64, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_float, c_void_p,
                              c_int, c_void_p, C.c_size_t, c_void_p]),
    "sa_gemm_nt": (c_int, [c_int64, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, C.POINTER(GemmEpilogue),
                           c_int64, c_void_p]),
    "sa_gemm_tn": (c_int, [c_int64, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_float, c_void_p,
                           c_int, c_void_p]),
    "sa_gemm_tn_colsum": (c_int, [c_int64, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_float,
                                  c_void_p, c_int, c_void_p, c_void_p]),
    "sa_embed_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, C.POINTER(c_void_p), c_void_p, c_int, c_int, c_int, c_int,
                             c_void_p, c_void_p, c_int, c_void_p]),
    "sa_embed_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, C.POINTER(C.c_int32), c_void_p,
                             C.POINTER(c_void_p), c_void_p, c_void_p]),
    "sa_favor_kmax": (c_int, [C.POINTER(FavorDesc), c_void_p, c_void_p, c_void_p, c_void_p]),
    "sa_favor_featmap_fwd": (c_int, [C.POINTER(FavorDesc), c_void_p, c_void_p, c_int, c_void_p, c_float, c_void_p, c_void_p,
                                     c_void_p]),
    "sa_favor_featmap_bwd": (c_int, [C.POINTER(FavorDesc), c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "sa_favor_kmax_fixup": (c_int, [C.POINTER(FavorDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sa_favor_scan_workspace": (C.c_size_t, [C.POINTER(FavorDesc), c_int]),
    "sa_favor_scan_fwd": (c_int, [C.POINTER(FavorDesc), c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p,
                                  c_void_p, C.c_size_t, c_void_p]),
    "sa_favor_scan_bwd": (c_int, [C.POINTER(FavorDesc), c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_size_t, c_void_p]),
    "sa_conv1x1_bwd_fused": (c_int, [c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "sa_conv1x1_bwd_fused_dbh": (c_int, [c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p]),
    "sa_conv1x1_fwd_fused": (c_int, [c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                     c_void_p]),
    "sa_favor_scan_states_bytes": (C.c_size_t, [C.POINTER(FavorDesc)]),
    "sa_favor_scan_fwd_save": (c_int, [C.POINTER(FavorDesc), c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int,
                                       c_void_p, c_void_p, C.c_size_t, c_void_p, C.c_size_t, c_void_p]),
    "sa_favor_scan_bwd_saved": (c_int, [C.POINTER(FavorDesc), c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                        c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_size_t, c_void_p,
                                        C.c_size_t, c_void_p]),
    "sa_local_attn_fwd": (c_int, [C.POINTER(LocalDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "sa_local_attn_bwd": (c_int, [C.POINTER(LocalDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "sa_local_attn_bwd_ws": (c_int, [C.POINTER(LocalDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sa_favor_scan_bwd_fused_supported": (c_int, [C.POINTER(FavorDesc)]),
    "sa_favor_scan_bwd_fused": (c_int, [C.POINTER(FavorDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_float, c_float, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, C.c_size_t, c_void_p, C.c_size_t, c_void_p]),
    "sa_rotary_table": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sa_local_attn_bwd_rot": (c_int, [C.POINTER(LocalDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sa_favor_decode_step": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int, c_void_p]),
    "sa_local_decode_step": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                     c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "sa_embed_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                              c_void_p, c_void_p, c_int, c_void_p]),
    "sa_tokens_prepare": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "sa_tokens_gather": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "sa_tokens_narrow": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "sa_rotary_qk": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "sa_rotary": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "sa_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p]),
    "sa_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "sa_ce_fwd_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sa_cast2d": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p]),
    "sa_gate_wgrad": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "sa_rezero_finish": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
}

_lock = threading.Lock()
_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) the CUDA library.  Raises if that is impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.library_path()
        if not os.path.exists(path):
            if not build_if_missing:
                raise RuntimeError(f"{path} is missing: run `python -m synthanatomy_b200.build` (needs nvcc). "
                                   "synthanatomy_b200 has no CPU / eager fallback.")
            path = _build.build_library()
        elif not _build.is_current():
            # the sources (csrc/, include/) changed since this .so was built: rebuild (under the cross-process build
            # lock) rather than run stale kernels against new headers
            if not build_if_missing:
                raise RuntimeError(f"{path} is stale (sources changed): run `python -m synthanatomy_b200.build`")
            path = _build.build_library()
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError here == ABI drift; let it propagate
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sa_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"synthanatomy_b200 {what} failed (status {rc}): {msg}")
