"""ctypes binding of the C ABI declared in include/pdes_b200.h.

The product path has NO fallback: if the nvcc-built sm_100a library is missing, importing the ops raises.
(`bind()` is also used by the CPU tests to attach the same prototypes to the emulation build of the very
same kernel sources; that library is test infrastructure and is never looked up here.)
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpdes_b200.so")

PDES_OK, PDES_ERR_ARG, PDES_ERR_UNSUPPORTED, PDES_ERR_LAUNCH = 0, 1, 2, 3
ACT_NONE, ACT_GELU = 0, 1

_P = c_void_p
_I = c_int

_PROTOTYPES = {
    # name: (restype, [argtypes])
    "pdes_version": (c_int, []),
    "pdes_last_error": (c_char_p, []),
    "pdes_is_cuda_build": (c_int, []),
    "pdes_tables_floats": (c_size_t, [_I, _I, _I, _I]),
    "pdes_tables_fill": (c_int, [_I, _I, _I, _I, _P]),
    "pdes_dft_fwd": (c_int, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P]),
    "pdes_mix_suggest_splits": (c_int, [_I, _I, _I, _I, _I]),
    "pdes_mix_fwd": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "pdes_mix_dx": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "pdes_mix_dw": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "pdes_mix_tc_ok": (c_int, [_I, _I, _I, _I, _I]),
    "pdes_mix_tc_pack_floats": (c_size_t, [_I, _I, _I, _I]),
    "pdes_mix_tc_x2_floats": (c_size_t, [_I, _I, _I, _I]),
    "pdes_mix_tc_o2_floats": (c_size_t, [_I, _I, _I, _I]),
    "pdes_mix_tc_pack": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "pdes_dft_fwd2": (c_int, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P]),
    "pdes_mix_tc_fwd": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "pdes_inv_h_modes": (c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "pdes_inv_h": (c_int, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "pdes_inv_w_gemm": (c_int, [_P, _P, _I, _P, _I, _P, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "pdes_set_tensor_core_mode": (None, [_I]),
    "pdes_get_tensor_core_mode": (c_int, []),
    "pdes_gemm_tc_supported": (c_int, [_I, _I]),
    "pdes_inv_w_gemm_tc_ok": (c_int, [_I, _I, _I, _I, _I, _P, _P]),
    "pdes_gemm_tc_pack_floats": (c_size_t, [_I, _I]),
    "pdes_gemm_tc_pack": (c_int, [_P, _I, _I, _I, _P, _P]),
    "pdes_gemm_tc_pack_t": (c_int, [_P, _I, _I, _I, _P, _P]),
    "pdes_inv_w_gemm_tc": (c_int, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "pdes_wgrad_tc_workspace_floats": (c_size_t, [_I, _I]),
    "pdes_wgrad_tc": (c_int, [_P, _P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _P]),
    "pdes_timeconv_ok": (c_int, [_I, _I]),
    "pdes_timeconv_forward": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "pdes_timeconv_bwd_workspace_floats": (c_size_t, [_I, _I, _I]),
    "pdes_timeconv_backward": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "pdes_constrain_forward": (c_int, [_P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "pdes_conv1x1_tc_ok": (c_int, [_I, _I, _I, _I, _P]),
    "pdes_conv1x1_tc": (c_int, [_P, _I, _P, _P, _P, _P, c_size_t, _I, _I, _I, _I, _P]),
    "pdes_wgrad_tc_range": (c_int, [_P, _P, _I, _I, _I, _P, _I, _P, _P, _I, _I, _I, _P]),
    "pdes_conv3x3_tc_pack_floats": (c_size_t, [_I, _I]),
    "pdes_conv3x3_tc_ok": (c_int, [_I, _I, _I, _I, _I, _P]),
    "pdes_conv3x3_tc": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "pdes_act_bwd": (c_int, [_P, _P, _P, c_size_t, _I, _P]),
    "pdes_transpose": (c_int, [_P, _P, _I, _I, _P]),
    "pdes_wgrad_workspace_floats": (c_size_t, [_I, _I, _I, _I]),
    "pdes_wgrad": (c_int, [_P, _P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _P]),
    "pdes_gn_workspace_bytes": (c_size_t, [_I, _I, _I, _I]),
    "pdes_gn_act_forward": (c_int, [_P, _P, _P, ctypes.c_float, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "pdes_gn_act_backward": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "pdes_block_fwd_workspace_floats": (c_size_t, [_I, _I, _I, _I, _I, _I, _I]),
    "pdes_block_forward": (c_int, [_P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                   _I, _I, _I, _I, _I, _I, _I, _P]),
    "pdes_block_bwd_workspace_floats": (c_size_t, [_I, _I, _I, _I, _I, _I, _I, _I]),
    "pdes_block_backward": (c_int, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                    _I, _I, _I, _I, _I, _I, _I, _P]),
    "pdes_mix_tc_dx_ok": (c_int, [_I, _I, _I, _I, _I, _I]),
    "pdes_mix_tc_dx": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


def bind(lib: ctypes.CDLL) -> ctypes.CDLL:
    """Attach restype/argtypes for every symbol of include/pdes_b200.h (raises AttributeError if one is missing)."""
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def check(lib: ctypes.CDLL, code: int) -> None:
    """Map a PDES_ERR_* code to the exception the reference would raise at the same point
    (assert / ValueError for bad modes or shapes, proc_fno.py:135-139; RuntimeError for device faults)."""
    if code == PDES_OK:
        return
    msg = lib.pdes_last_error().decode("utf-8", "replace")
    if code == PDES_ERR_ARG:
        raise ValueError(msg)
    if code == PDES_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


_lib = None


def library() -> ctypes.CDLL:
    """The sm_100a library.  Fails loudly when it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or ./build.sh).  neural_pde_surrogates_b200 has no CPU or PyTorch fallback for the spectral block.")
        lib = bind(ctypes.CDLL(LIB_PATH))
        if lib.pdes_is_cuda_build() != 1:
            raise RuntimeError(f"{LIB_PATH} is not an nvcc/sm_100a build")
        _lib = lib
    return _lib
