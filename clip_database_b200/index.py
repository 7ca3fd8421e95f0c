"""``GpuIndex`` — one GPU's resident float32 row store behind the C ABI.

Thin, typed wrapper over ``include/clipdb.h``; all arithmetic happens in the
CUDA library.  Host-array methods are synchronous (like the reference's
``cursor.execute`` / ``fetchall`` pair, image_database.py:1582-1583); the
``*_device`` methods take torch CUDA tensors, enqueue on the context's stream
and do not synchronise.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib

METRICS = {"cosine": _lib.METRIC_COSINE, "l2": _lib.METRIC_L2}


@dataclass
class SearchResult:
    """Per query: ``rowids[q, :counts[q]]`` / ``distances[q, :counts[q]]`` sorted by
    (distance ascending, rowid ascending); ``nan_rows[q]`` admitted rows whose
    distance was NaN (excluded from the result, see clipdb.h)."""
    rowids: np.ndarray      # int64 [nq, k]
    distances: np.ndarray   # float32 [nq, k]
    counts: np.ndarray      # int32 [nq]
    nan_rows: np.ndarray    # int64 [nq]

    def row(self, q: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        m = int(self.counts[q])
        return self.rowids[q, :m], self.distances[q, :m]


def _metric(metric) -> int:
    if isinstance(metric, str):
        try:
            return METRICS[metric.lower()]
        except KeyError:
            raise ValueError(f"unknown metric {metric!r}; expected one of {sorted(METRICS)}")
    return int(metric)


def _fptr(a: Optional[np.ndarray]):
    if a is None:
        return ctypes.cast(None, ctypes.POINTER(ctypes.c_float))
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class GpuIndex:
    def __init__(self, device: int = 0):
        self._L = _lib.load()
        self._ctx = ctypes.c_void_p()
        rc = self._L.clipdb_create(int(device), ctypes.byref(self._ctx))
        if rc != _lib.OK:
            self._ctx = ctypes.c_void_p()
            raise _lib.ClipdbError(rc, "clipdb_create failed (no usable CUDA device? there is no CPU fallback)")
        self.device = int(device)
        self._keepalive = []  # tensors borrowed by attach()
        self.batch_enabled = False

    # ---- lifetime -----------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.clipdb_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()
            self._keepalive = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        _lib.check(self._ctx, rc)

    # ---- plumbing -----------------------------------------------------------------
    @property
    def num_rows(self) -> int:
        return int(self._L.clipdb_num_rows(self._ctx))

    @property
    def dim(self) -> int:
        return int(self._L.clipdb_dim(self._ctx))

    @property
    def launch_count(self) -> int:
        return int(self._L.clipdb_launch_count(self._ctx))

    def set_option(self, name: str, value: int) -> None:
        self._check(self._L.clipdb_set_option(self._ctx, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = ctypes.c_int64(0)
        self._check(self._L.clipdb_get_option(self._ctx, name.encode(), ctypes.byref(v)))
        return int(v.value)

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        """Run on the given ``cudaStream_t`` (an int handle; 0 = CUDA's legacy default stream);
        ``None`` = back to the context's own stream."""
        if cuda_stream is None:
            self._check(self._L.clipdb_use_own_stream(self._ctx))
        else:
            self._check(self._L.clipdb_set_stream(self._ctx, ctypes.c_void_p(int(cuda_stream))))

    def use_torch_stream(self) -> None:
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self) -> None:
        self._check(self._L.clipdb_synchronize(self._ctx))

    def profile(self, enable: bool) -> None:
        """Bracket every scan kernel with CUDA events on the launching stream."""
        self._check(self._L.clipdb_profile(self._ctx, int(bool(enable))))

    def profile_read(self) -> Tuple[float, int]:
        """(summed scan-kernel milliseconds, number of scans) since ``profile(True)``."""
        ms = ctypes.c_double(0.0)
        n = ctypes.c_int64(0)
        self._check(self._L.clipdb_profile_read(self._ctx, ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    def profile_clock(self) -> Tuple[Optional[float], int]:
        """(SM clock in MHz the profiled contraction launches of the batched path ran at — %clock64 over
        %globaltimer inside the kernel —, launches) since the last call."""
        cyc, ns, n = ctypes.c_double(0.0), ctypes.c_double(0.0), ctypes.c_int64(0)
        self._check(self._L.clipdb_profile_clock(self._ctx, ctypes.byref(cyc), ctypes.byref(ns), ctypes.byref(n)))
        return (cyc.value / ns.value * 1e3 if ns.value > 0 else None), int(n.value)

    # ---- row store ------------------------------------------------------------------
    def load(self, rows, rowids=None) -> None:
        """Copy ``rows`` (numpy ``[n, dim]`` float32, or a torch tensor on any device)
        into context-owned HBM.  ``rowids`` (int64, ascending = scan order) optional."""
        if _is_torch(rows):
            rows_t = rows.detach().contiguous().float()
            n, dim = rows_t.shape
            ids_t = None
            if rowids is not None:
                import torch
                ids_t = torch.as_tensor(rowids, dtype=torch.int64).contiguous()
            self._check(self._L.clipdb_load_rows(self._ctx, ctypes.c_void_p(rows_t.data_ptr()),
                                                 ctypes.c_void_p(ids_t.data_ptr() if ids_t is not None else 0),
                                                 n, dim))
            return
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2:
            raise ValueError("rows must be [n, dim]")
        n, dim = rows.shape
        ids = None
        if rowids is not None:
            ids = np.ascontiguousarray(rowids, dtype=np.int64)
            if ids.shape != (n,):
                raise ValueError("rowids must be [n]")
        self._check(self._L.clipdb_load_rows(self._ctx, ctypes.c_void_p(rows.ctypes.data),
                                             ctypes.c_void_p(ids.ctypes.data if ids is not None else 0),
                                             n, dim))
        self._keepalive = []

    def reserve(self, capacity: int, dim: int, explicit_rowids: bool = True, placement: str = "device",
                device_rows: int = 0) -> None:
        """Empty store with room for ``capacity`` rows, filled by ``append`` chunk by chunk (a streaming
        loader never holds more than a chunk on the host).  ``placement="host"`` keeps the float32 rows
        in pinned, device-mapped host memory — all of them, or those from position ``device_rows`` on
        (rounded down to a multiple of 128; the rows before stay in HBM): combined with ``enable_batch``
        and ``set_option("batch_min_nq", 1)`` that is a bf16-primary store (see clipdb_reserve_rows)."""
        places = {"device": _lib.PLACE_DEVICE, "host": _lib.PLACE_HOST}
        if placement not in places:
            raise ValueError(f"placement must be one of {sorted(places)}")
        if placement == "host":
            self.set_option("host_tier_from_row", int(device_rows))
        self._check(self._L.clipdb_reserve_rows(self._ctx, int(capacity), int(dim), int(bool(explicit_rowids)),
                                                places[placement]))
        self._keepalive = []
        self.batch_enabled = False

    def stage_buffer(self, rows: int, dim: int) -> np.ndarray:
        """A float32 ``[rows, dim]`` numpy view of context-owned PINNED host memory: fill it, then
        ``append(view[:m], ...)`` moves the rows with one DMA instead of a pageable-memory copy.  Valid
        until the next call that asks for a larger buffer, or ``close``."""
        ptr = ctypes.c_void_p()
        nbytes = int(rows) * int(dim) * 4
        self._check(self._L.clipdb_stage_buffer(self._ctx, nbytes, ctypes.byref(ptr)))
        buf = (ctypes.c_float * (int(rows) * int(dim))).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.float32).reshape(int(rows), int(dim))

    def append_sqlite(self, db_path: str, min_rowid: Optional[int] = None, max_rowid: Optional[int] = None,
                      on_chunk=None, chunk_rows: int = 8192) -> Tuple[int, int]:
        """Native loader (clipdb_append_sqlite): append the joined vec0 rows with rowid in (min_rowid, max_rowid] of
        a reference-schema SQLite file to the reserved store, read by SQLite's C library straight into pinned
        memory.  ``on_chunk(rowids, image_ids, last_modified, file_paths)`` receives numpy copies and the list of
        paths per chunk.  Returns (vec0 rows in the range before the joins, rows appended).  Raises ``ClipdbError``
        with ``code == _lib.ERR_UNSUPPORTED`` when vec0 is not a plain table or libsqlite3 is unavailable."""
        errors = []

        def trampoline(_user, n, p_ids, p_images, p_mtimes, p_paths, paths_bytes):
            try:
                n = int(n)
                ids = np.ctypeslib.as_array(p_ids, shape=(n,)).copy()
                images = np.ctypeslib.as_array(p_images, shape=(n,)).copy()
                mtimes = np.ctypeslib.as_array(p_mtimes, shape=(n,)).copy()
                blob = ctypes.string_at(p_paths, int(paths_bytes))
                paths = blob.decode("utf-8").split("\0")[:n]
                if on_chunk is not None:
                    on_chunk(ids, images, mtimes, paths)
                return 0
            except BaseException as e:      # noqa: BLE001 — an exception must not unwind through the C frames
                errors.append(e)
                return 1
        cb = _lib.SQLITE_CHUNK_FN(trampoline)
        vec0_rows, joined = ctypes.c_int64(0), ctypes.c_int64(0)
        lo = -(1 << 63) if min_rowid is None else int(min_rowid)
        hi = (1 << 63) - 1 if max_rowid is None else int(max_rowid)
        rc = self._L.clipdb_append_sqlite(self._ctx, os.fsencode(db_path), lo, hi, int(chunk_rows),
                                          ctypes.cast(cb, ctypes.c_void_p), None, ctypes.byref(vec0_rows),
                                          ctypes.byref(joined))
        if errors:
            raise errors[0]
        self._check(rc)
        return int(vec0_rows.value), int(joined.value)

    def append(self, rows, rowids=None) -> None:
        """Append rows (numpy, or a torch tensor on any device) in scan order."""
        if _is_torch(rows):
            rows_t = rows.detach().contiguous()
            if rows_t.dtype.is_floating_point is False or rows_t.element_size() != 4 or rows_t.dim() != 2:
                raise ValueError("rows must be a float32 [m, dim] tensor")
            ids_t = None
            if rowids is not None:
                import torch
                ids_t = torch.as_tensor(rowids, dtype=torch.int64).contiguous()
            self._check(self._L.clipdb_append_rows(self._ctx, ctypes.c_void_p(rows_t.data_ptr()),
                                                   ctypes.c_void_p(ids_t.data_ptr() if ids_t is not None else 0),
                                                   rows_t.shape[0]))
            return
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or (self.dim and rows.shape[1] != self.dim):
            raise ValueError("rows must be [m, dim] with the store's dim")
        ids = None
        if rowids is not None:
            ids = np.ascontiguousarray(rowids, dtype=np.int64)
        self._check(self._L.clipdb_append_rows(self._ctx, ctypes.c_void_p(rows.ctypes.data),
                                               ctypes.c_void_p(ids.ctypes.data if ids is not None else 0),
                                               rows.shape[0]))

    def update_row(self, position: int, row) -> None:
        row = np.ascontiguousarray(row, dtype=np.float32)
        if row.shape != (self.dim,):
            raise ValueError("row must be [dim]")
        self._check(self._L.clipdb_update_row(self._ctx, int(position), ctypes.c_void_p(row.ctypes.data)))

    def attach(self, rows, rowids=None, rowid_base: int = 0) -> None:
        """Borrow a torch CUDA tensor ``[n, dim]`` float32 (no copy).  Kept alive here."""
        if not (_is_torch(rows) and rows.is_cuda and rows.is_contiguous() and rows.dtype.is_floating_point
                and rows.element_size() == 4):
            raise ValueError("attach() needs a contiguous float32 CUDA tensor")
        n, dim = rows.shape
        ids_ptr = 0
        if rowids is not None:
            if not (rowids.is_cuda and rowids.is_contiguous() and rowids.element_size() == 8):
                raise ValueError("rowids must be a contiguous int64 CUDA tensor")
            ids_ptr = rowids.data_ptr()
        self._check(self._L.clipdb_attach_rows(self._ctx, ctypes.c_void_p(rows.data_ptr()),
                                               ctypes.c_void_p(ids_ptr), n, dim, int(rowid_base)))
        self._keepalive = [rows, rowids]

    def set_mask(self, admitted) -> None:
        """``admitted``: bool/uint8 array of length ``num_rows``; True = row takes part."""
        bits = np.asarray(admitted).astype(bool)
        if bits.shape != (self.num_rows,):
            raise ValueError("mask must have one entry per row")
        packed = np.packbits(bits, bitorder="little")
        words = np.zeros((self.num_rows + 31) // 32, dtype=np.uint32)
        words.view(np.uint8)[: packed.shape[0]] = packed
        self._check(self._L.clipdb_set_mask(self._ctx, ctypes.c_void_p(words.ctypes.data), words.shape[0]))

    def set_mask_words(self, words) -> None:
        """The bitset itself: ``ceil(num_rows / 32)`` 32-bit words (bit ``r & 31`` of word ``r >> 5``
        = row r admitted) as a numpy array or a torch tensor on any device; copied."""
        need = (self.num_rows + 31) // 32
        if _is_torch(words):
            if words.element_size() != 4 or not words.is_contiguous() or words.numel() < need:
                raise ValueError("mask words must be a contiguous 32-bit tensor with one bit per row")
            self._check(self._L.clipdb_set_mask(self._ctx, ctypes.c_void_p(words.data_ptr()), words.numel()))
            return
        w = np.ascontiguousarray(words, dtype=np.uint32)
        if w.shape[0] < need:
            raise ValueError("mask words must hold one bit per row")
        self._check(self._L.clipdb_set_mask(self._ctx, ctypes.c_void_p(w.ctypes.data), w.shape[0]))

    def clear_mask(self) -> None:
        self._check(self._L.clipdb_clear_mask(self._ctx))

    # ---- sign codes (the reference's binary fallback search) -------------------------------------
    @property
    def num_codes(self) -> int:
        return int(self._L.clipdb_num_codes(self._ctx))

    def load_codes(self, codes, ids=None) -> None:
        """``codes``: uint8 ``[n, 1152]`` of 0/1 (numpy, or a torch tensor on any device), as stored
        in ``binary_embeddings`` (image_database.py:1189-1190); bit-packed on the GPU.  ``ids``
        (int64 ``[n]``) are what ``binary_search`` returns instead of scan positions."""
        if _is_torch(codes):
            t = codes.detach().contiguous()
            if t.element_size() != 1 or t.dim() != 2:
                raise ValueError("codes must be [n, dim] uint8")
            n, dim = t.shape
            ptr = t.data_ptr()
        else:
            t = np.ascontiguousarray(codes, dtype=np.uint8)
            if t.ndim != 2:
                raise ValueError("codes must be [n, dim] uint8")
            n, dim = t.shape
            ptr = t.ctypes.data
        ids_a = None
        if ids is not None:
            ids_a = np.ascontiguousarray(ids, dtype=np.int64)
            if ids_a.shape != (n,):
                raise ValueError("ids must be [n]")
        self._check(self._L.clipdb_load_codes(self._ctx, ctypes.c_void_p(ptr),
                                              ctypes.c_void_p(ids_a.ctypes.data if ids_a is not None else 0), n, dim))

    def set_code_mask(self, admitted, order=None) -> None:
        """Admission bitset of the code store (bool per code) and, optionally, the tie-break
        sequence of every code while the mask is in use (see clipdb_set_code_mask)."""
        n = self.num_codes
        bits = np.asarray(admitted).astype(bool)
        if bits.shape != (n,):
            raise ValueError("mask must have one entry per code")
        packed = np.packbits(bits, bitorder="little")
        words = np.zeros((n + 31) // 32, dtype=np.uint32)
        words.view(np.uint8)[: packed.shape[0]] = packed
        order_a = None
        if order is not None:
            order_a = np.ascontiguousarray(order, dtype=np.uint32)
            if order_a.shape != (n,):
                raise ValueError("order must have one entry per code")
        self._check(self._L.clipdb_set_code_mask(self._ctx, ctypes.c_void_p(words.ctypes.data), words.shape[0],
                                                 ctypes.c_void_p(order_a.ctypes.data if order_a is not None else 0)))

    def clear_code_mask(self) -> None:
        self._check(self._L.clipdb_clear_code_mask(self._ctx))

    def binary_search(self, query_code, k: int, score_mode: str = "reference", use_mask: bool = False
                      ) -> Tuple[np.ndarray, np.ndarray]:
        """Top-k codes by AND-popcount with ``query_code`` (uint8 ``[1152]`` of 0/1).  Returns
        (ids int64, scores int32), score descending, ties in scan order.  ``score_mode``:
        "reference" = modulo 256 like the reference's uint8 ``np.dot``; "popcount" = plain."""
        modes = {"reference": _lib.SCORE_REFERENCE_UINT8, "popcount": _lib.SCORE_POPCOUNT}
        if score_mode not in modes:
            raise ValueError(f"score_mode must be one of {sorted(modes)}")
        q = np.ascontiguousarray(query_code, dtype=np.uint8).ravel()
        if q.shape[0] != 1152:
            raise ValueError("query code must be uint8[1152]")
        kc = max(int(k), 0)
        ids = np.full(max(kc, 1), -1, dtype=np.int64)
        scores = np.zeros(max(kc, 1), dtype=np.int32)
        n = ctypes.c_int32(0)
        self._check(self._L.clipdb_binary_search(self._ctx, ctypes.c_void_p(q.ctypes.data), kc, modes[score_mode],
                                                 int(bool(use_mask)), ctypes.c_void_p(ids.ctypes.data),
                                                 ctypes.c_void_p(scores.ctypes.data), ctypes.byref(n)))
        return ids[:n.value].copy(), scores[:n.value].copy()

    def binary_search_device(self, d_query_words, k: int, out_ids, out_scores, out_n,
                             score_mode: str = "reference", use_mask: bool = False) -> None:
        """Async: torch CUDA tensors (query: 36 packed 32-bit words; outputs int64 ``[k]``, int32 ``[k]``,
        int32 ``[1]``), enqueued on the context's stream."""
        modes = {"reference": _lib.SCORE_REFERENCE_UINT8, "popcount": _lib.SCORE_POPCOUNT}
        self._check(self._L.clipdb_binary_search_device(
            self._ctx, ctypes.c_void_p(d_query_words.data_ptr()), int(k), modes[score_mode], int(bool(use_mask)),
            ctypes.c_void_p(out_ids.data_ptr()), ctypes.c_void_p(out_scores.data_ptr()),
            ctypes.c_void_p(out_n.data_ptr())))

    # ---- query arithmetic -----------------------------------------------------------------
    @staticmethod
    def _pack_negatives(negatives, negative_weights, dim):
        negs = [np.ascontiguousarray(v, dtype=np.float32).reshape(dim) for v in negatives]
        ws = [float(w) for w in negative_weights]
        if len(negs) != len(ws):
            raise ValueError("one weight per negative")
        if not negs:
            return None, None, 0
        return np.ascontiguousarray(np.stack(negs)), (ctypes.c_double * len(ws))(*ws), len(negs)

    def blend(self, e1, e2=None, weights: Tuple[float, float] = (0.5, 0.5),
              negatives: Sequence = (), negative_weights: Sequence[float] = ()
              ) -> Tuple[np.ndarray, int]:
        """K3 on the GPU.  Returns (float32[dim], CLIPDB_BLEND_* flags)."""
        e1 = np.ascontiguousarray(e1, dtype=np.float32).ravel()
        dim = e1.shape[0]
        e2a = None if e2 is None else np.ascontiguousarray(e2, dtype=np.float32).reshape(dim)
        negs, ws, n_neg = self._pack_negatives(negatives, negative_weights, dim)
        out = np.empty(dim, dtype=np.float32)
        flags = ctypes.c_int32(0)
        self._check(self._L.clipdb_blend(self._ctx, _fptr(e1), _fptr(e2a), float(weights[0]),
                                         float(weights[1]), _fptr(negs), ws, n_neg, dim, _fptr(out),
                                         ctypes.byref(flags)))
        return out, int(flags.value)

    # ---- search ---------------------------------------------------------------------------
    def _alloc_result(self, nq: int, k: int) -> SearchResult:
        kc = max(k, 0)
        return SearchResult(np.full((nq, kc), -1, dtype=np.int64),
                            np.full((nq, kc), np.nan, dtype=np.float32),
                            np.zeros(nq, dtype=np.int32), np.zeros(nq, dtype=np.int64))

    def search(self, queries, k: int, metric="cosine", use_mask: bool = False) -> SearchResult:
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [nq, {self.dim}]")
        nq = q.shape[0]
        res = self._alloc_result(nq, int(k))
        self._check(self._L.clipdb_search(
            self._ctx, _fptr(q), nq, int(k), _metric(metric), int(bool(use_mask)),
            res.rowids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _fptr(res.distances),
            res.counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            res.nan_rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))
        return res

    def blend_search(self, e1, k: int, e2=None, weights: Tuple[float, float] = (0.5, 0.5),
                     negatives: Sequence = (), negative_weights: Sequence[float] = (),
                     metric="cosine", use_mask: bool = False, return_query: bool = False):
        """blend + scan + top-k in one host call; the blended query stays in HBM."""
        e1 = np.ascontiguousarray(e1, dtype=np.float32).ravel()
        dim = e1.shape[0]
        if dim != self.dim:
            raise ValueError(f"query must be [{self.dim}]")
        e2a = None if e2 is None else np.ascontiguousarray(e2, dtype=np.float32).reshape(dim)
        negs, ws, n_neg = self._pack_negatives(negatives, negative_weights, dim)
        res = self._alloc_result(1, int(k))
        qout = np.empty(dim, dtype=np.float32) if return_query else None
        flags = ctypes.c_int32(0)
        self._check(self._L.clipdb_blend_search(
            self._ctx, _fptr(e1), _fptr(e2a), float(weights[0]), float(weights[1]), _fptr(negs), ws,
            n_neg, int(k), _metric(metric), int(bool(use_mask)),
            res.rowids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _fptr(res.distances),
            res.counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            res.nan_rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _fptr(qout),
            ctypes.byref(flags)))
        if return_query:
            return res, qout, int(flags.value)
        return res

    def search_device(self, d_queries, k: int, out_rowids, out_dist, out_n, out_nan=None,
                      metric="cosine", use_mask: bool = False) -> None:
        """Async: torch CUDA tensors in and out (queries ``[nq, dim]`` float32; outputs
        int64 ``[nq, k]``, float32 ``[nq, k]``, int32 ``[nq]``, int64 ``[nq]``)."""
        nq = d_queries.shape[0] if d_queries.dim() == 2 else 1
        self._check(self._L.clipdb_search_device(
            self._ctx, ctypes.c_void_p(d_queries.data_ptr()), nq, int(k), _metric(metric),
            int(bool(use_mask)), ctypes.c_void_p(out_rowids.data_ptr()),
            ctypes.c_void_p(out_dist.data_ptr()), ctypes.c_void_p(out_n.data_ptr()),
            ctypes.c_void_p(out_nan.data_ptr() if out_nan is not None else 0)))

    def enable_batch(self, enable: bool = True) -> None:
        """Build (or drop) the bf16 copy of the store used by the tensor-core batched path.
        Afterwards ``search()`` with 2 or more queries uses it; results are identical."""
        self._check(self._L.clipdb_enable_batch(self._ctx, int(bool(enable))))
        self.batch_enabled = bool(enable)

    def batch_stats(self) -> Tuple[np.ndarray, np.ndarray]:
        """(candidates kept by the tensor-core filter, candidates re-ranked in float32) per query
        slot of the last batched pass (uint32 [256] each)."""
        cand = np.zeros(256, dtype=np.uint32)
        surv = np.zeros(256, dtype=np.uint32)
        self._check(self._L.clipdb_batch_stats(self._ctx, ctypes.c_void_p(cand.ctypes.data),
                                               ctypes.c_void_p(surv.ctypes.data)))
        return cand, surv

    def search_ptrs(self, q_ptr: int, nq: int, k: int, p_rowids: int, p_dist: int, p_n: int, p_nan: int,
                    metric="cosine", use_mask: bool = False) -> None:
        """Async exact search on raw device addresses (outputs nq x k row-major)."""
        self._check(self._L.clipdb_search_device(
            self._ctx, ctypes.c_void_p(q_ptr), nq, int(k), _metric(metric), int(bool(use_mask)),
            ctypes.c_void_p(p_rowids), ctypes.c_void_p(p_dist), ctypes.c_void_p(p_n), ctypes.c_void_p(p_nan)))

    def search_batch_ptrs(self, q_ptr: int, nq: int, k: int, p_rowids: int, p_dist: int, p_n: int, p_nan: int,
                          p_flags: int, use_mask: bool = False) -> None:
        """Async batched (tensor-core) search on raw device addresses, nq <= 256."""
        self._check(self._L.clipdb_search_batch_device(
            self._ctx, ctypes.c_void_p(q_ptr), nq, int(k), int(bool(use_mask)), ctypes.c_void_p(p_rowids),
            ctypes.c_void_p(p_dist),
            ctypes.c_void_p(p_n), ctypes.c_void_p(p_nan), ctypes.c_void_p(p_flags)))

    def merge_batch_records_device(self, records, nq: int, k: int, off_rowids: int, off_dist: int, off_count: int,
                                   out_dist, out_rowids, out_n) -> None:
        """Async merge of gathered per-rank BATCH records ``[lists, record_bytes]`` (uint8 CUDA tensor):
        within a record, query q's arrays sit q*k*8 / q*k*4 / q*4 bytes after the given offsets."""
        lists, rec = records.shape
        base = records.data_ptr()
        self._check(self._L.clipdb_merge_batch_device(
            self._ctx, ctypes.c_void_p(base + off_dist), rec, 4 * k, ctypes.c_void_p(base + off_rowids), rec, 8 * k,
            ctypes.c_void_p(base + off_count), rec, 4, lists, int(nq), int(k),
            ctypes.c_void_p(out_dist.data_ptr()), ctypes.c_void_p(out_rowids.data_ptr()),
            ctypes.c_void_p(out_n.data_ptr())))

    def search_batch_device(self, d_queries, k: int, out_rowids, out_dist, out_n, out_nan, flags,
                            use_mask: bool = False) -> None:
        """Async batched search (<= 256 queries): torch CUDA tensors; ``flags[q] != 0`` marks
        queries that must be re-run with ``search_device`` (see clipdb.h)."""
        nq = d_queries.shape[0]
        self._check(self._L.clipdb_search_batch_device(
            self._ctx, ctypes.c_void_p(d_queries.data_ptr()), nq, int(k), int(bool(use_mask)),
            ctypes.c_void_p(out_rowids.data_ptr()),
            ctypes.c_void_p(out_dist.data_ptr()), ctypes.c_void_p(out_n.data_ptr()),
            ctypes.c_void_p(out_nan.data_ptr() if out_nan is not None else 0), ctypes.c_void_p(flags.data_ptr())))

    def blend_device(self, d_e1, d_e2, d_w, d_negs, d_neg_w, d_out, d_flags=None) -> None:
        """Async batched K3: see clipdb_blend_device."""
        batch, dim = d_e1.shape
        n_neg = 0 if d_negs is None else d_negs.shape[1]
        ptr = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
        self._check(self._L.clipdb_blend_device(self._ctx, ptr(d_e1), ptr(d_e2), ptr(d_w), ptr(d_negs),
                                                ptr(d_neg_w), n_neg, dim, batch, ptr(d_out), ptr(d_flags)))

    def merge_records_device(self, records, k: int, off_rowids: int, off_dist: int, off_count: int,
                             out_dist, out_rowids, out_n) -> None:
        """Async shard merge of ``records``: a uint8 CUDA tensor ``[lists, record_bytes]`` holding
        one packed per-rank result each (see ``sharded.RecordLayout``)."""
        lists, rec = records.shape
        base = records.data_ptr()
        self._check(self._L.clipdb_merge_strided_device(
            self._ctx, ctypes.c_void_p(base + off_dist), rec, ctypes.c_void_p(base + off_rowids), rec,
            ctypes.c_void_p(base + off_count), rec, lists, int(k), ctypes.c_void_p(out_dist.data_ptr()),
            ctypes.c_void_p(out_rowids.data_ptr()), ctypes.c_void_p(out_n.data_ptr())))

    def search_into_record(self, d_query, k: int, record, off_rowids: int, off_dist: int, off_count: int,
                           off_nan: int, metric="cosine", use_mask: bool = False) -> None:
        """Async single-query search whose outputs land inside ``record`` (uint8 CUDA tensor)."""
        base = record.data_ptr()
        self._check(self._L.clipdb_search_device(
            self._ctx, ctypes.c_void_p(d_query.data_ptr()), 1, int(k), _metric(metric), int(bool(use_mask)),
            ctypes.c_void_p(base + off_rowids), ctypes.c_void_p(base + off_dist),
            ctypes.c_void_p(base + off_count), ctypes.c_void_p(base + off_nan)))

    # ---- fused shard exchange (one launch per sharded query, no collective call) ---------------
    def exchange_init(self, world: int, rank: int) -> Tuple[bytes, int]:
        """Allocate this rank's inbox.  Returns (CUDA IPC handle bytes for the other processes,
        raw device pointer for ranks living in this process)."""
        handle = (ctypes.c_uint8 * _lib.IPC_HANDLE_BYTES)()
        ptr = ctypes.c_void_p()
        self._check(self._L.clipdb_exchange_init(self._ctx, int(world), int(rank), handle, ctypes.byref(ptr)))
        return bytes(handle), int(ptr.value)

    def exchange_connect(self, handles: Sequence[bytes]) -> None:
        """``handles``: every rank's IPC handle in rank order (one process per GPU)."""
        blob = b"".join(handles)
        buf = (ctypes.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._check(self._L.clipdb_exchange_connect(self._ctx, buf))

    def exchange_connect_pointers(self, inboxes: Sequence[int], devices: Optional[Sequence[int]] = None) -> None:
        """Same-process variant: every rank's inbox pointer (and CUDA device) in rank order."""
        ptrs = (ctypes.c_void_p * len(inboxes))(*[ctypes.c_void_p(int(p)) for p in inboxes])
        devs = (ctypes.c_int32 * len(inboxes))(*[int(d) for d in devices]) if devices is not None else None
        self._check(self._L.clipdb_exchange_connect_pointers(self._ctx, ptrs, devs))

    def exchange_abort(self, abort: bool = True) -> None:
        """Make every exchange of this context that is waiting for its peers give up now (out_n = -1)
        instead of spinning until ``xchg_timeout_ms``; ``False`` re-arms.  Takes no lock: callable from
        a watchdog thread while another thread waits on the stream."""
        rc = self._L.clipdb_exchange_abort(self._ctx, int(bool(abort)))
        if rc != _lib.OK:
            raise _lib.ClipdbError(rc, "exchange_abort: exchange not initialised")

    def exchange_set_epoch(self, single_epoch: int, batch_epoch: int = 0) -> None:
        """Next single-query / batched exchange uses sequence number ``single_epoch + 1`` /
        ``batch_epoch + 1``.  Every rank must be idle and pass the same values (resynchronisation
        after a rank failed to enqueue a search)."""
        self._check(self._L.clipdb_exchange_set_epoch(self._ctx, int(single_epoch) & 0xFFFFFFFF,
                                                      int(batch_epoch) & 0xFFFFFFFF))

    def exchange_stats(self, enable: bool = True, reset: bool = True):
        """({"scan", "local_merge", "publish", "wait_peers", "final_merge"} -> mean microseconds per
        launch, launches) of the fused sharded search since the last reset; ``enable`` switches the
        in-kernel %globaltimer stamping on or off from the next launch."""
        ns = (ctypes.c_double * 5)()
        n = ctypes.c_int64(0)
        self._check(self._L.clipdb_exchange_stats(self._ctx, int(bool(enable)), int(bool(reset)), ns, ctypes.byref(n)))
        names = ("scan", "local_merge", "publish", "wait_peers", "final_merge")
        m = max(int(n.value), 1)
        return {k: ns[i] / m / 1e3 for i, k in enumerate(names)}, int(n.value)

    def search_sharded_device(self, d_query, k: int, out_rowids, out_dist, out_n, out_nan=None,
                              metric="cosine", use_mask: bool = False) -> None:
        """Async: this shard's scan + exchange with the peer GPUs + merge in ONE launch; the outputs
        (torch CUDA tensors, int64 ``[k]``, float32 ``[k]``, int32 ``[1]``, int64 ``[1]``) hold the
        answer over ALL shards.  ``out_n == -1``: a peer did not deliver in time."""
        self._check(self._L.clipdb_search_sharded_device(
            self._ctx, ctypes.c_void_p(d_query.data_ptr()), int(k), _metric(metric), int(bool(use_mask)),
            ctypes.c_void_p(out_rowids.data_ptr()), ctypes.c_void_p(out_dist.data_ptr()),
            ctypes.c_void_p(out_n.data_ptr()), ctypes.c_void_p(out_nan.data_ptr() if out_nan is not None else 0)))

    def search_batch_sharded_device(self, d_queries, k: int, out_rowids, out_dist, out_n, out_nan, flags,
                                    use_mask: bool = False) -> None:
        """Async batched search over ALL shards (<= 256 queries; ``enable_batch`` on every rank): the
        batched path's last kernel exchanges each query's candidates with the peer GPUs and merges.
        ``flags[q] != 0``: some shard could not answer query q through the batched path."""
        nq = d_queries.shape[0]
        self._check(self._L.clipdb_search_batch_sharded_device(
            self._ctx, ctypes.c_void_p(d_queries.data_ptr()), nq, int(k), int(bool(use_mask)),
            ctypes.c_void_p(out_rowids.data_ptr()), ctypes.c_void_p(out_dist.data_ptr()),
            ctypes.c_void_p(out_n.data_ptr()), ctypes.c_void_p(out_nan.data_ptr() if out_nan is not None else 0),
            ctypes.c_void_p(flags.data_ptr())))

    def merge_device(self, d_dist, d_rowids, d_counts, k: int, out_dist, out_rowids, out_n) -> None:
        """Async shard merge of ``[lists, k]`` gathered results (clipdb_merge_device)."""
        lists = d_dist.shape[0]
        self._check(self._L.clipdb_merge_device(
            self._ctx, ctypes.c_void_p(d_dist.data_ptr()), ctypes.c_void_p(d_rowids.data_ptr()),
            ctypes.c_void_p(d_counts.data_ptr()), lists, int(k), ctypes.c_void_p(out_dist.data_ptr()),
            ctypes.c_void_p(out_rowids.data_ptr()), ctypes.c_void_p(out_n.data_ptr())))
