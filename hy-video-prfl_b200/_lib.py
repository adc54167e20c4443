"""ctypes binding of libprfl_b200.so (the C ABI declared in include/prfl_b200.h).

There is deliberately no fallback: if the shared library is missing or a kernel returns an
error, this module raises.  Nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprfl_b200.so")

PRFL_OK = 0
EPI_BF16, EPI_BF16_GELU, EPI_F32, EPI_RESIDUAL, EPI_BF16_DGELU = 0, 1, 2, 3, 4


class PrflError(RuntimeError):
    pass


_lib = None

_i64, _i32, _f32, _p = C.c_int64, C.c_int, C.c_float, C.c_void_p

_SIGS = {
    "prfl_abi_version": (C.c_int, []),
    "prfl_last_error_string": (C.c_char_p, []),
    "prfl_launch_count": (_i64, []),
    "prfl_launch_count_reset": (None, []),
    "prfl_ln_mod_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _f32, _i32, _p]),
    "prfl_ln_mod_split_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _f32, _p]),
    "prfl_rmsnorm_rope_fwd": (C.c_int, [_p, _i64, _p, _p, _p, _p, _i64, _p, _i64, _i32, _i64, _i64, _f32, _p]),
    "prfl_gemm_bf16": (C.c_int, [_p, _i64, _i32, _p, _i64, _i32, _p, _i64, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p]),
    "prfl_attn_fwd_ws_bytes": (_i64, [_i32, _i32, _i32]),
    "prfl_attn_fwd_split_plan": (None, [_i32, _i32, _i32, _i32, _p]),
    "prfl_attn_fwd": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i32, _i32, _i32, _f32, _p, _p]),
    "prfl_attn_merge": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _i32, _p, _i64, _i64, _i32, _i32, _p]),
    "prfl_colsum_parts": (C.c_int, [_i64]),
    "prfl_ln_mod_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _p]),
    "prfl_rmsnorm_rope_bwd": (C.c_int, [_p, _i64, _p, _p, _p, _p, _i64, _p, _p, _i64, _p, _i64, _i64, _i32, _i64, _i64, _p]),
    "prfl_colsum_bf16": (C.c_int, [_p, _i64, _p, _i64, _i32, _p]),
    "prfl_gate_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _p]),
    "prfl_attn_bwd_ws_floats": (_i64, [_i32, _i32]),
    "prfl_attn_bwd": (C.c_int, [_p, _i64, _i64] * 5 + [_p, _p] + [_p, _i64, _i64] * 3 + [_i32, _i32, _i32, _f32, _p]),
    "prfl_attn_fwd_p2p": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i32, _i32, _i32, _i64, _i64, _p, _i32, _i32,
                                    _i32, _f32, _p, _p]),
    "prfl_a2a_scatter_p2p": (C.c_int, [_p, _i64, _i64, _p, _i32, _i32, _i32, _i32, _p]),
    "prfl_a2a_gather_p2p": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _i32, _i32, _i32, _i32, _p]),
    "prfl_patchify": (C.c_int, [_p, _i32, _p, _i32, _p, _i32, _i32, _i32, _p]),
    "prfl_patchify_bwd": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _i32, _p]),
    "prfl_unpatchify": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "prfl_sq_pool_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _p]),
    "prfl_sq_pool_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "prfl_cast_f32_bf16": (C.c_int, [_p, _p, _i64, _p]),
    "prfl_unipc_step": (C.c_int, [_p, _p, _p, _f32, _p, _p, _p, _p, _f32, _p, _p, _p, _p, _p, _i64, _p]),
    "prfl_scale2_f32": (C.c_int, [_p, _f32, _p, _f32, _p, _i64, _p]),
    "prfl_adamw_step": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _p]),
    "prfl_sumsq_f32": (C.c_int, [_p, _i64, _p, _p]),
    "prfl_a2a_pack": (C.c_int, [_p, _i64, _i64, _p, _i32, _i32, _i32, _i32, _p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PrflError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(prfl_b200 has no CPU / PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != PRFL_OK:
        msg = lib().prfl_last_error_string()
        raise PrflError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(lib().prfl_launch_count())


def launch_count_reset():
    lib().prfl_launch_count_reset()
