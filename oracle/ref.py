"""ctypes face of ``oracle_ref.c`` plus a numpy fp64 ground truth.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED.

Three tiers:
  * ``exact_knn``      — fp64 accumulate, lexicographic (distance, sequence) sort.
  * ``knn``            — the C restatement of the scalar float32 loop + SQLite's
                         bounded ORDER BY/LIMIT admission rule (oracle_ref.c).
  * ``sql_harness``    — real SQLite running the reference's statement.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
LIB_PATH = os.path.join(_BUILD, "liboracle_ref.so")
SHIM_PATH = os.path.join(_BUILD, "vec_shim.so")

_lib = None


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc only)."""
    if force or not (os.path.exists(LIB_PATH) and os.path.exists(SHIM_PATH)):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []),
                       check=True, capture_output=True)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        i64p = ctypes.POINTER(ctypes.c_int64)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.oracle_cosine_distance_f32.restype = ctypes.c_float
        L.oracle_cosine_distance_f32.argtypes = [f32p, f32p, ctypes.c_int64]
        L.oracle_distances.restype = None
        L.oracle_distances.argtypes = [f32p, ctypes.c_int64, ctypes.c_int64, f32p, f32p]
        L.oracle_distances_mt.restype = None
        L.oracle_distances_mt.argtypes = [f32p, ctypes.c_int64, ctypes.c_int64, f32p, f32p,
                                          ctypes.c_int]
        L.oracle_distances_f64.restype = None
        L.oracle_distances_f64.argtypes = [f32p, ctypes.c_int64, ctypes.c_int64, f32p, f64p]
        L.oracle_knn.restype = ctypes.c_int64
        L.oracle_knn.argtypes = [f32p, i64p, ctypes.c_int64, ctypes.c_int64, f32p, u8p,
                                 ctypes.c_int64, i64p, f32p, i64p, i64p]
        L.oracle_fill_unit_rows.restype = None
        L.oracle_fill_unit_rows.argtypes = [f32p, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_uint64]
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return ctypes.cast(None, ctypes.POINTER(ctype))
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def cosine_distance(a, b) -> float:
    """One ``vec_distance_cosine(a, b)`` as SQLite would see it (float32 value)."""
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape and a.ndim == 1
    return float(lib().oracle_cosine_distance_f32(_ptr(a, ctypes.c_float),
                                                  _ptr(b, ctypes.c_float), a.shape[0]))


def distances(rows, query, threads: int = 1) -> np.ndarray:
    rows, query = _f32(rows), _f32(query)
    n, d = rows.shape
    out = np.empty(n, dtype=np.float32)
    if threads <= 1:
        lib().oracle_distances(_ptr(rows, ctypes.c_float), n, d, _ptr(query, ctypes.c_float),
                               _ptr(out, ctypes.c_float))
    else:
        lib().oracle_distances_mt(_ptr(rows, ctypes.c_float), n, d,
                                  _ptr(query, ctypes.c_float), _ptr(out, ctypes.c_float),
                                  threads)
    return out


def distances_f64(rows, query) -> np.ndarray:
    rows, query = _f32(rows), _f32(query)
    n, d = rows.shape
    out = np.empty(n, dtype=np.float64)
    lib().oracle_distances_f64(_ptr(rows, ctypes.c_float), n, d, _ptr(query, ctypes.c_float),
                               _ptr(out, ctypes.c_double))
    return out


def knn(rows, query, k: int, rowids=None, mask=None
        ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """``oracle_ref``: (rowids, float32 distances, scan sequence, n_nan).

    Restates the statement at image_database.py:1564-1574 (see oracle_ref.c).
    """
    rows, query = _f32(rows), _f32(query)
    n, d = rows.shape
    assert query.shape == (d,)
    if rowids is not None:
        rowids = np.ascontiguousarray(rowids, dtype=np.int64)
        assert rowids.shape == (n,)
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        assert mask.shape == (n,)
    cap = n if (k < 0 or k > n) else k
    o_id = np.empty(max(cap, 1), dtype=np.int64)
    o_d = np.empty(max(cap, 1), dtype=np.float32)
    o_s = np.empty(max(cap, 1), dtype=np.int64)
    nn = ctypes.c_int64(0)
    m = lib().oracle_knn(_ptr(rows, ctypes.c_float), _ptr(rowids, ctypes.c_int64), n, d,
                         _ptr(query, ctypes.c_float), _ptr(mask, ctypes.c_uint8), k,
                         _ptr(o_id, ctypes.c_int64), _ptr(o_d, ctypes.c_float),
                         _ptr(o_s, ctypes.c_int64), ctypes.byref(nn))
    if m < 0:
        raise MemoryError("oracle_knn")
    return o_id[:m].copy(), o_d[:m].copy(), o_s[:m].copy(), int(nn.value)


def exact_knn(rows, query, k: int, rowids=None, mask=None
              ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``oracle_exact``: fp64 distances, stable (distance, sequence) order."""
    d = distances_f64(rows, query)
    seq = np.arange(d.shape[0], dtype=np.int64)
    keep = ~np.isnan(d)
    if mask is not None:
        keep &= np.asarray(mask).astype(bool)
    seq = seq[keep]
    order = seq[np.argsort(d[keep], kind="stable")]
    if k >= 0:
        order = order[:k]
    ids = order if rowids is None else np.asarray(rowids, dtype=np.int64)[order]
    return ids, d[order], order


def fill_unit_rows(n: int, dim: int, seed: int) -> np.ndarray:
    """Cheap deterministic approximately-normal unit rows (bench sample data)."""
    out = np.empty((n, dim), dtype=np.float32)
    lib().oracle_fill_unit_rows(_ptr(out, ctypes.c_float), n, dim, seed)
    return out
