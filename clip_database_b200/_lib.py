"""ctypes binding of ``include/clipdb.h`` (the C-ABI shared library).

There is no fallback: if ``libclipdb_b200.so`` is missing, or a call returns a
non-zero code, this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32,
                    c_void_p)

from . import build as _build

OK = 0
ERR_NAMES = {1: "INVALID", 2: "CUDA", 3: "NOMEM", 4: "STATE", 5: "UNSUPPORTED"}
METRIC_COSINE = 0
METRIC_L2 = 1
SCORE_REFERENCE_UINT8 = 0   # popcount(q AND row) modulo 256, as numpy's uint8 dot product gives
SCORE_POPCOUNT = 1
IPC_HANDLE_BYTES = 64
BLEND_POSITIVE_ZERO_NORM = 1
BLEND_NEGATIVE_ZERO_NORM = 2
ABI_VERSION = 5
PLACE_DEVICE = 0
PLACE_HOST = 1

# every symbol include/clipdb.h declares: (name, restype, argtypes)
_F = POINTER(c_float)
_I64 = POINTER(c_int64)
_I32 = POINTER(c_int32)
_U32 = POINTER(c_uint32)
_D = POINTER(c_double)
_CTX = c_void_p

# clipdb_sqlite_chunk_fn: (user, n, rowids, image_ids, last_modified, paths, paths_bytes) -> int
SQLITE_CHUNK_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int64, POINTER(c_int64), POINTER(c_int64), POINTER(c_double),
                                   POINTER(ctypes.c_char), c_int64)
ERR_UNSUPPORTED = 5

SIGNATURES = [
    ("clipdb_abi_version", c_int, []),
    ("clipdb_source_hash", c_char_p, []),
    ("clipdb_create", c_int, [c_int, POINTER(_CTX)]),
    ("clipdb_destroy", None, [_CTX]),
    ("clipdb_last_error", c_char_p, [_CTX]),
    ("clipdb_set_stream", c_int, [_CTX, c_void_p]),
    ("clipdb_use_own_stream", c_int, [_CTX]),
    ("clipdb_synchronize", c_int, [_CTX]),
    ("clipdb_set_option", c_int, [_CTX, c_char_p, c_int64]),
    ("clipdb_get_option", c_int, [_CTX, c_char_p, _I64]),
    ("clipdb_launch_count", c_int64, [_CTX]),
    ("clipdb_load_rows", c_int, [_CTX, c_void_p, c_void_p, c_int64, c_int32]),
    ("clipdb_append_rows", c_int, [_CTX, c_void_p, c_void_p, c_int64]),
    ("clipdb_update_row", c_int, [_CTX, c_int64, c_void_p]),
    ("clipdb_attach_rows", c_int, [_CTX, c_void_p, c_void_p, c_int64, c_int32, c_int64]),
    ("clipdb_reserve_rows", c_int, [_CTX, c_int64, c_int32, c_int32, c_int32]),
    ("clipdb_stage_buffer", c_int, [_CTX, c_int64, POINTER(c_void_p)]),
    ("clipdb_append_sqlite", c_int, [_CTX, c_char_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, _I64, _I64]),
    ("clipdb_num_rows", c_int64, [_CTX]),
    ("clipdb_dim", c_int32, [_CTX]),
    ("clipdb_set_mask", c_int, [_CTX, c_void_p, c_int64]),
    ("clipdb_clear_mask", c_int, [_CTX]),
    ("clipdb_blend", c_int, [_CTX, _F, _F, c_double, c_double, _F, _D, c_int32, c_int32, _F, _I32]),
    ("clipdb_blend_device", c_int, [_CTX, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    ("clipdb_search", c_int, [_CTX, _F, c_int32, c_int32, c_int32, c_int32, _I64, _F, _I32, _I64]),
    ("clipdb_search_device", c_int, [_CTX, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    ("clipdb_blend_search", c_int, [_CTX, _F, _F, c_double, c_double, _F, _D, c_int32, c_int32,
                                    c_int32, c_int32, _I64, _F, _I32, _I64, _F, _I32]),
    ("clipdb_enable_batch", c_int, [_CTX, c_int32]),
    ("clipdb_search_batch_device", c_int, [_CTX, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p]),
    ("clipdb_batch_stats", c_int, [_CTX, c_void_p, c_void_p]),
    ("clipdb_load_codes", c_int, [_CTX, c_void_p, c_void_p, c_int64, c_int32]),
    ("clipdb_num_codes", c_int64, [_CTX]),
    ("clipdb_set_code_mask", c_int, [_CTX, c_void_p, c_int64, c_void_p]),
    ("clipdb_clear_code_mask", c_int, [_CTX]),
    ("clipdb_binary_search", c_int, [_CTX, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    ("clipdb_binary_search_device", c_int, [_CTX, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                            c_void_p]),
    ("clipdb_exchange_init", c_int, [_CTX, c_int32, c_int32, c_void_p, c_void_p]),
    ("clipdb_exchange_connect", c_int, [_CTX, c_void_p]),
    ("clipdb_exchange_connect_pointers", c_int, [_CTX, c_void_p, c_void_p]),
    ("clipdb_search_sharded_device", c_int, [_CTX, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                             c_void_p]),
    ("clipdb_search_batch_sharded_device", c_int, [_CTX, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                                   c_void_p, c_void_p, c_void_p]),
    ("clipdb_exchange_abort", c_int, [_CTX, c_int32]),
    ("clipdb_exchange_set_epoch", c_int, [_CTX, c_uint32, c_uint32]),
    ("clipdb_exchange_stats", c_int, [_CTX, c_int32, c_int32, _D, _I64]),
    ("clipdb_merge_device", c_int, [_CTX, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                    c_void_p, c_void_p, c_void_p]),
    ("clipdb_merge_strided_device", c_int, [_CTX, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                            c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    ("clipdb_merge_batch_device", c_int, [_CTX, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                          c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    ("clipdb_profile", c_int, [_CTX, c_int32]),
    ("clipdb_profile_clock", c_int, [_CTX, _D, _D, _I64]),
    ("clipdb_profile_read", c_int, [_CTX, _D, _I64]),
]

_lib = None


class ClipdbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"clipdb error {code} ({ERR_NAMES.get(code, '?')}): {message}")
        self.code = code


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    """dlopen the in-tree CUDA library and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if _build.is_stale():
        # missing, or compiled from other sources than the ones next to it (the library is not under
        # version control): never run old kernels silently — rebuild when a compiler is here, else refuse
        try:
            _build.build()
        except RuntimeError as e:
            raise RuntimeError(
                f"{path} is missing or was not built from the current sources, and it cannot be rebuilt "
                f"here ({e}).  There is no CPU fallback: run `python -m clip_database_b200.build` where "
                "nvcc is available.") from e
    L = ctypes.CDLL(path)
    for name, restype, argtypes in SIGNATURES:
        fn = getattr(L, name)  # AttributeError = symbol not exported
        fn.restype = restype
        fn.argtypes = argtypes
    got = L.clipdb_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"{path}: ABI version {got}, binding expects {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(ctx, rc: int) -> None:
    if rc != OK:
        msg = load().clipdb_last_error(ctx)
        raise ClipdbError(rc, msg.decode("utf-8", "replace") if msg else "")
