"""ctypes binding of libatmvfi_b200.so (C ABI declared in include/atmvfi.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no fallback: if
the shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libatmvfi_b200.so")

MAX_SRC = 4
FP32, TF32, TF32X3, F16 = 0, 1, 2, 3
OUT_PIXEL, OUT_SHUFFLE2, OUT_WINDOW_REV, OUT_QKV_HEADS = 0, 1, 2, 3


class AtmvfiError(RuntimeError):
    pass


class WindowGeom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B2", "H", "W", "ws", "shift", "Hp", "Wp", "pad_top", "pad_left")]


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int32), ("pitch", C.c_int32)]


class GemmConvDesc(C.Structure):
    _fields_ = [
        ("nsrc", C.c_int32),
        ("src", Src * MAX_SRC),
        ("B", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32), ("dil", C.c_int32),
        ("Hout", C.c_int32), ("Wout", C.c_int32), ("Cout", C.c_int32),
        ("weight", C.c_void_p), ("ldw", C.c_int32),
        ("bias", C.c_void_p), ("prelu", C.c_void_p),
        ("residual", C.c_void_p), ("res_pitch", C.c_int32),
        ("out", C.c_void_p), ("out_pitch", C.c_int32),
        ("out2", C.c_void_p), ("prelu2", C.c_void_p), ("out2_pitch", C.c_int32),
        ("out_mode", C.c_int32),
        ("win", WindowGeom),
        ("precision", C.c_int32),
        ("tma_host", C.c_void_p),
        ("row_begin", C.c_int32), ("row_end", C.c_int32),
        ("qkv_heads", C.c_int32),
        ("out_f32", C.c_int32),
        ("head32", C.c_void_p), ("head32_pitch", C.c_int32), ("head32_c0", C.c_int32),
        ("pad_stores", C.c_int32), ("param_pad", C.c_int32),
    ]


class P2PPiece(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("chunk_bytes", C.c_uint64), ("chunk_stride", C.c_uint64),
                ("nchunks", C.c_uint32), ("reserved", C.c_uint32)]


IPC_HANDLE_BYTES, P2P_MAX_PIECES, P2P_MAX_PEERS = 64, 16, 8


class RowOwners(C.Structure):
    _fields_ = [("nseg", C.c_int32), ("row_lo", C.c_int32 * (P2P_MAX_PEERS * 2 + 1)), ("byte_delta", C.c_int64 * (P2P_MAX_PEERS * 2))]


_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_int64
_GP, _DP = C.POINTER(WindowGeom), C.POINTER(GemmConvDesc)

# name -> argtypes; every function returns int (0 = ok) unless listed in _SPECIAL
PROTOTYPES = {
    "atmvfi_gemm_conv": [_DP, _P],
    "atmvfi_gemm_conv_plan": [_DP, _P],
    "atmvfi_layernorm": [_P, _I, _P, _I, _L, _I, _P, _P, _F, _P],
    # ... every grid walker ends with (y0, y1, stream): the row window of include/atmvfi.h
    "atmvfi_window_gather_ln": [_P, _I, _P, _I, _I, _GP, _P, _P, _F, _I, _I, _P],
    "atmvfi_window_attention": [_P, _I, _P, _I, _I, _I, _GP, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _P],
    "atmvfi_conv3x3_first": [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "atmvfi_pack5_planar": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "atmvfi_window_attention_tc": [_P, _I, _P, _I, _I, _I, _GP, _I, _P, _I, _P, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _P],
    "atmvfi_dwconv3x3_gelu": [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P],
    "atmvfi_mlp_tail": [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "atmvfi_flow_warp_nchw": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "atmvfi_flow_warp_nhwc": [_P, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "atmvfi_flow_warp_nhwc_p2p": [_P, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, C.POINTER(RowOwners), _P],
    "atmvfi_warp_blend": [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "atmvfi_warp_blend_p2p": [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, C.POINTER(RowOwners), _P],
    "atmvfi_resize_bilinear_ac": [_P, _P, _I, _I, _I, _I, _I, _F, _I, _I, _P],
    "atmvfi_pyramid_warp": [_P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "atmvfi_nchw_to_nhwc": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "atmvfi_nhwc_to_nchw": [_P, _I, _P, _I, _I, _I, _I, _P],
    "atmvfi_l1_mean": [_P, _P, _P, _P, _I, _L, _P],
    "atmvfi_select_min3": [_P, _P, _P, _P, _P, _P, _P, _I, _L, _P],
    "atmvfi_copy": [_P, _P, C.c_size_t, _P],
    "atmvfi_residual_finish": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "atmvfi_u8_to_planar": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "atmvfi_planar_to_u8": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "atmvfi_u8_to_planar_rows": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    # NVLink row exchange (spatial row-slab mode)
    "atmvfi_arena_alloc": [C.c_size_t, C.POINTER(C.c_void_p)],
    "atmvfi_arena_free": [_P],
    "atmvfi_ipc_export": [_P, C.c_char_p],
    "atmvfi_ipc_open": [C.c_char_p, C.POINTER(C.c_void_p)],
    "atmvfi_ipc_close": [_P],
    "atmvfi_p2p_exchange": [C.POINTER(P2PPiece), _I, C.POINTER(C.c_void_p), _I, C.POINTER(C.c_void_p), _I, _P, _P, _P, _P],
    "atmvfi_p2p_step_begin": [_P, C.POINTER(C.c_void_p), _I, C.POINTER(C.c_void_p), _I, _P, _P],
    "atmvfi_p2p_set_timeout_ms": [_I],
    "atmvfi_cast_f32_to_f16": [_P, _I, _P, _I, _L, _I, _I, _P],
}
_SPECIAL = {
    "atmvfi_last_error": ([], C.c_char_p),
    "atmvfi_abi_version": ([], C.c_int),
    "atmvfi_device_info": ([_I, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
    "atmvfi_gemm_conv_plan_bytes": ([], C.c_int),
    "atmvfi_l1_mean_scratch_floats": ([C.c_int], C.c_int),
    "atmvfi_attn_prof_read": ([C.POINTER(C.c_uint64)], C.c_int),
    "atmvfi_mlp_tail_prof_read": ([C.POINTER(C.c_uint64)], C.c_int),
    "atmvfi_set_output_rounding": ([C.c_int], None),
    "atmvfi_set_activation_f16": ([C.c_int], None),
}
ALL_SYMBOLS = sorted(list(PROTOTYPES) + list(_SPECIAL))

_lib = None


def load() -> C.CDLL:
    """Load the shared object (once).  Raises AtmvfiError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AtmvfiError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU or PyTorch fallback for the ATM-VFI kernels.")
    lib = C.CDLL(LIB_PATH)
    for name, argt in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = argt, C.c_int
    for name, (argt, rest) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = argt, rest
    if lib.atmvfi_abi_version() != 3:
        raise AtmvfiError("libatmvfi_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().atmvfi_last_error().decode("utf-8", "replace")
        raise AtmvfiError(f"{what} failed (status {status}): {msg}")
