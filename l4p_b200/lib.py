"""ctypes binding of the C-ABI CUDA library `libl4p_b200.so` (declared in include/l4p_b200.h).

There is no CPU fallback: if the library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

_PKG = Path(__file__).resolve().parent
# L4P_LIB selects an experiment build of the same ABI (tools/ A/B runs only; see l4p_b200/build.py L4P_BUILD_TAG)
LIB_PATH = Path(os.environ["L4P_LIB"]).resolve() if os.environ.get("L4P_LIB") else _PKG / "libl4p_b200.so"

ACT_NONE, ACT_GELU, ACT_RELU, ACT_EXP = 0, 1, 2, 3
STORE_ROWMAJOR, STORE_QKV, STORE_CONVT, STORE_HEAD1X1, STORE_HYPER = 0, 1, 2, 3, 4
A_MATRIX, A_CONV3D = 0, 1


class L4PError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    """Mirror of `l4p_gemm_desc` (include/l4p_b200.h)."""

    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("lda", C.c_int64), ("ldw", C.c_int64),
        ("bf16", C.c_int), ("a_mode", C.c_int),
        ("cB", C.c_int), ("cT", C.c_int), ("cH", C.c_int), ("cW", C.c_int), ("cCin", C.c_int),
        ("kT", C.c_int), ("kH", C.c_int), ("kW", C.c_int),
        ("bT", C.c_int), ("bH", C.c_int), ("bW", C.c_int),
        ("bias", C.c_void_p), ("act", C.c_int),
        ("res_f32", C.c_void_p), ("res_16", C.c_void_p), ("res2_16", C.c_void_p), ("ld_res", C.c_int64),
        ("res_row_mod", C.c_int),
        ("store_mode", C.c_int),
        ("out_f32", C.c_void_p), ("out_16", C.c_void_p), ("out_16_relu", C.c_void_p), ("ld_out", C.c_int64),
        ("q", C.c_void_p), ("k", C.c_void_p), ("vt", C.c_void_p),
        ("heads", C.c_int), ("head_dim", C.c_int), ("head_dim_pad", C.c_int), ("tokens", C.c_int),
        ("sT", C.c_int), ("sH", C.c_int), ("sW", C.c_int), ("ctCout", C.c_int),
        ("w2", C.c_void_p), ("b2", C.c_void_p), ("c2", C.c_int), ("exp_out", C.c_int),
        ("rows_per_group", C.c_int64),
        ("block_n", C.c_int),
        ("cta_pair", C.c_int),
        ("prof", C.c_void_p),
        ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_int64), ("split_k", C.c_int),
        ("grp_a_rows", C.c_int64), ("grp_b_rows", C.c_int64), ("m_stride", C.c_int), ("conv_grp_b", C.c_int),
    ]


# name -> (restype, argtypes); every symbol include/l4p_b200.h declares must be listed here
# (tests/test_abi.py cross-checks this table against the header).
_SIGNATURES = {
    "l4p_version": (C.c_int, []),
    "l4p_last_error": (C.c_char_p, []),
    "l4p_init": (C.c_int, [C.c_int, C.c_int]),
    "l4p_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                C.c_float, C.c_int, C.c_void_p]),
    "l4p_gemm": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "l4p_gemm_plan": (C.c_int, [C.POINTER(GemmDesc), C.POINTER(C.c_int)]),
    "l4p_patchify": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]),
    "l4p_preprocess_rgb": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 11 + [C.c_void_p, C.c_void_p, C.c_void_p]),
    "l4p_cast16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "l4p_cast16_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "l4p_upsample3d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 10 + [C.c_void_p]),
    "l4p_im2col3": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]),
    "l4p_pose_from_rays": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_float, C.c_int] + [C.c_void_p] * 6),
    "l4p_affine_align_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "l4p_affine_align_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "l4p_token_attention": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_int64, C.c_float, C.c_int, C.c_void_p]),
    "l4p_image_attention": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_float, C.c_int, C.c_void_p]),
    "l4p_layernorm16": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p]),
    "l4p_head_expand": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]),
    "l4p_head_diag_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "l4p_row_softmax16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "l4p_group_softmax_t16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "l4p_token_weighted_sum": (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "l4p_track_readout": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 7 + [C.c_void_p]),
    "l4p_sim3_align": (C.c_int, [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p]),
    "l4p_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def exported_symbols():
    return sorted(_SIGNATURES)


def load(build_if_missing: bool = False) -> C.CDLL:
    """Load the shared library and bind all prototypes. Raises L4PError if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if build_if_missing:
            from . import build as _b

            _b.build()
        else:
            raise L4PError(
                f"{LIB_PATH} not found: build it with `python -m l4p_b200.build` "
                "(there is no CPU fallback for the l4p_b200 hot path)"
            )
    lib = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().l4p_last_error().decode(errors="replace")
        raise L4PError(f"{what or 'l4p call'} failed ({rc}): {msg}")
