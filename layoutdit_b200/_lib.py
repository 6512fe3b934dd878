"""ctypes binding of libldit_b200.so (include/ldit.h).  There is no fallback: if the
library is missing or a call fails, the caller gets an exception."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDIT_LIB_PATH") or os.path.join(_HERE, "libldit_b200.so")  # override: A/B builds in experiments

_c = ctypes
_vp, _i, _f = _c.c_void_p, _c.c_int, _c.c_float

# name -> (restype, argtypes); mirrors include/ldit.h one to one
SIGNATURES = {
    "ldit_version": (_i, []),
    "ldit_error_string": (_c.c_char_p, [_i]),
    "ldit_layernorm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "ldit_gemm_bias": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ldit_gemm_bias_gelu": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ldit_gemm_bias_scale_residual": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ldit_gemm_accumulate": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "ldit_gemm_bias_scale": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ldit_add_layernorm": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "ldit_transpose_bf16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "ldit_colsum_bf16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "ldit_gelu": (_i, [_vp, _vp, _c.c_size_t, _vp]),
    "ldit_gelu_bwd": (_i, [_vp, _vp, _vp, _c.c_size_t, _vp]),
    "ldit_scale_residual": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "ldit_scale_residual_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "ldit_scale_residual_rows": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _vp]),
    "ldit_scale_residual_rows_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "ldit_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "ldit_gemm_wgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "ldit_gemm_dgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "ldit_attention_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "ldit_attention_lse": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ldit_attention_bwd_flash": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ldit_resample_taps_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "ldit_batch_sum": (_i, [_vp, _vp, _i, _i, _vp]),
    "ldit_mlp_clusters": (_i, []),
    "ldit_mlp_schedule": (_i, [_i, _i, _i, _vp, _i]),
    "ldit_mlp_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "ldit_patch_embed": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ldit_patch_embed_tma": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ldit_patch_embed_tma_preferred": (_i, [_i, _i, _i, _i]),
    "ldit_patch_embed_pages": (_i, [_vp, _vp, _i, _i, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ldit_attention": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ldit_resample_taps": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "ldit_fpn_merge": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "ldit_conv3x3_bias": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ldit_conv3x3_bias_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ldit_subsample2": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ldit_subsample2_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ldit_resize_rows": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ldit_patch_embed_scratch_bytes": (_c.c_size_t, [_i, _i, _i]),
    "ldit_set_gemm_tile_n": (None, [_i]),
    "ldit_set_gemm_cta_pair": (None, [_i]),
    "ldit_set_attention_impl": (None, [_i]),
    "ldit_set_pdl": (None, [_i]),
    "ldit_set_l2_window": (_i, [_vp, _vp, _c.c_size_t, _c.c_size_t]),
    "ldit_workspace_bytes": (_c.c_size_t, [_i, _i, _i, _i, _i]),
    "ldit_has_experimental": (_i, []),
    "ldit_debug_attention_timeline": (None, [_vp]),
    "ldit_debug_gemm_timeline": (None, [_vp]),
    "ldit_launch_count": (_c.c_ulonglong, []),
    "ldit_reset_launch_count": (None, []),
}

DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2

_lib = None


class LditError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LditError(f"{LIB_PATH} is missing: run `python -m layoutdit_b200.build` "
                        "(or __graft_entry__.build()). There is no CPU or eager fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drifted apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().ldit_error_string(rc).decode()
        raise LditError(f"{what} failed: {msg} (code {rc})")
