"""ctypes binding of the C-ABI shared library (include/loco_asr.h).

There is NO CPU / PyTorch fallback: if ``libloco_asr.so`` is missing or a CUDA device is absent the
product path raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LOCO_ASR_LIB") or os.path.join(_HERE, "libloco_asr.so")   # env override: A/B builds in tools/
DEBUG_LIB_PATH = os.path.join(_HERE, "libloco_asr_debug.so")     # -DLOCO_DEBUG: product kernels + cross-check kernels + knobs (tests)
ABI_VERSION = 2
CSRC = os.path.join(_HERE, "csrc")

LOCO_F32, LOCO_F16, LOCO_BF16, LOCO_F64 = 0, 1, 2, 3
LOCO_OK, LOCO_ERR_INVALID, LOCO_ERR_CUDA, LOCO_ERR_WEIGHTS, LOCO_ERR_WORKSPACE, LOCO_ERR_STATE = 0, -1, -2, -3, -4, -5   # loco_status
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL = 0, 1, 2
EPI_LN_BIAS, EPI_LN_BIAS_GELU, EPI_BIAS_RESIDUAL_STATS, EPI_BIAS_LNRESIDUAL_STATS = 3, 4, 5, 6


class LocoError(RuntimeError):
    pass


class LocoConfigC(C.Structure):
    _fields_ = [
        ("hidden_size", C.c_int32), ("encoder_layers", C.c_int32), ("encoder_attention_heads", C.c_int32),
        ("encoder_ffn_dim", C.c_int32), ("num_conv_layers", C.c_int32),
        ("conv_dim", C.c_int32 * 8), ("conv_kernel", C.c_int32 * 8), ("conv_stride", C.c_int32 * 8),
        ("num_conv_pos_embeddings", C.c_int32), ("num_conv_pos_embedding_groups", C.c_int32),
        ("max_speech_positions", C.c_int32), ("encoder_max_relative_position", C.c_int32),
        ("pad_token_id", C.c_int32), ("feat_extract_norm_is_group", C.c_int32),
        ("activation_is_gelu", C.c_int32), ("conv_bias", C.c_int32), ("layer_norm_eps", C.c_float),
    ]


# name -> (restype, argtypes); every symbol include/loco_asr.h declares (tests check the export list)
_H = C.c_void_p
SIGNATURES = {
    "loco_abi_version": (C.c_int, []),
    "loco_default_config": (None, [C.POINTER(LocoConfigC)]),
    "loco_create": (C.c_int, [C.POINTER(LocoConfigC), C.c_int, C.POINTER(_H)]),
    "loco_destroy": (None, [_H]),
    "loco_last_error": (C.c_char_p, [_H]),
    "loco_load_tensor": (C.c_int, [_H, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "loco_finalize_weights": (C.c_int, [_H]),
    "loco_plan": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_size_t)]),
    "loco_encode": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "loco_host_workspace_bytes": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "loco_encode_host": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "loco_plan_create": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int, C.POINTER(_H)]),
    "loco_plan_info": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_size_t)]),
    "loco_encode_planned": (C.c_int, [_H, _H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "loco_plan_destroy": (None, [_H, _H]),
    "loco_sync_check": (C.c_int, [_H, C.c_void_p]),
    "loco_is_debug_build": (C.c_int, []),
    "loco_plan_text": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_size_t)]),
    "loco_encode_text": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "loco_set_head": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "loco_set_head_outputs": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "loco_launch_count": (C.c_int64, [_H]),
    "loco_profile_enable": (C.c_int, [_H, C.c_int]),
    "loco_profile_collect": (C.c_int, [_H, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "loco_debug_set": (C.c_int, [_H, C.c_char_p, C.c_int64]),
    "loco_debug_buffer": (C.c_int, [_H, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "loco_debug_gemm": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "loco_debug_gemm_ln": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_libs = {}


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a with nvcc (in-tree, so the .so files travel to the GPU box): the product library and
    its LOCO_DEBUG twin."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise LocoError("building libloco_asr.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def load(debug: bool = False):
    """dlopen the library and attach signatures.  Raises LocoError if it has not been built.  ``debug=True`` loads the
    LOCO_DEBUG build (cross-check kernels + the knobs that select them), which only the tests and tools/ use."""
    path = os.environ.get("LOCO_ASR_LIB") or (DEBUG_LIB_PATH if debug else LIB_PATH)     # the override serves both
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise LocoError(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(or `make -C loco_asr_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library drift
        fn.restype = res
        fn.argtypes = args
    if lib.loco_abi_version() != ABI_VERSION:
        raise LocoError(f"ABI version mismatch: library {lib.loco_abi_version()} != binding {ABI_VERSION}")
    if bool(lib.loco_is_debug_build()) != bool(debug) and not os.environ.get("LOCO_ASR_LIB"):
        raise LocoError(f"{path} is {'a' if lib.loco_is_debug_build() else 'not a'} LOCO_DEBUG build")
    _libs[path] = lib
    return lib


def check(lib, handle, rc: int, what: str):
    if rc != 0:
        msg = lib.loco_last_error(handle)
        raise LocoError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")
