"""``ImageDatabase`` — the reference's search surface on top of the GPU index.

Mirrors ``ImageDatabase.search`` (image_database.py:1308-1658): same signature,
same result type (``List[Tuple[file_path, similarity]]`` sorted by similarity
descending), same guards and fallbacks — with the body replaced: the store is
loaded once from the SQLite database into resident HBM and every search is one
blend + scan + top-k on the GPU instead of a per-row SQL function call.

The SigLIP model is out of scope (no weights offline): ``search()`` takes text /
image-path queries when an ``embedder`` is supplied, and ``vector:<file.npy>`` queries
(ready-made embeddings) always; ``search_embedding()`` takes the float32[1152] vectors
directly and is what ``search()`` calls after embedding.  There is no CPU search path.
"""
from __future__ import annotations

import os
import sqlite3
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib, loader, schema
from .index import GpuIndex

Result = List[Tuple[str, float]]


def _serialised(method):
    """One search / refresh / reload at a time per ``ImageDatabase``: a refresh from one thread (``dropin`` runs it
    before every search) must not move the host-side row metadata under a search in another."""
    import functools

    @functools.wraps(method)
    def locked(self, *args, **kwargs):
        with self._lock:
            return method(self, *args, **kwargs)
    return locked


class Embedder:
    """Interface for the out-of-scope model: text / image path -> float32[dim] (or None)."""

    def text(self, query: str) -> Optional[np.ndarray]:       # image_database.py:509-543
        raise NotImplementedError

    def image(self, path: str) -> Optional[np.ndarray]:       # image_database.py:443-463
        raise NotImplementedError


VECTOR_PREFIX = "vector:"


def load_vector(path: str, dim: int) -> np.ndarray:
    """A ready-made query embedding from a ``.npy`` file (``vector:<path>`` in any query position of
    ``search()`` and of the interactive session): float32[dim], e.g. saved from the reference's
    ``_get_text_embedding`` / ``_get_image_embedding`` or from any SigLIP 2 SO400M pipeline."""
    v = np.asarray(np.load(path, allow_pickle=False), dtype=np.float32).reshape(-1)
    if v.shape[0] != dim:
        raise ValueError(f"{path}: {v.shape[0]} values, expected {dim}")
    return v


def load_embedder(spec: str) -> Embedder:
    """``module:Class`` or ``module:factory`` -> an object with ``text(str)`` and ``image(path)`` methods
    returning float32[1152] (or None).  The class / factory is called without arguments."""
    import importlib
    module, _, name = spec.partition(":")
    if not module or not name:
        raise ValueError("embedder must be given as module:Class")
    obj = getattr(importlib.import_module(module), name)()
    for method in ("text", "image"):
        if not callable(getattr(obj, method, None)):
            raise TypeError(f"{spec}: the embedder needs a {method}() method")
    return obj


def like_prefix_mask(file_paths: Sequence[str], folders: Sequence[str],
                     lowered: Optional[List[bytes]] = None) -> np.ndarray:
    """Rows the reference's folder WHERE clause admits (image_database.py:1513-1529, 1576-1579).

    The reference normalises each folder with ``os.path.abspath`` + a trailing
    separator, escapes ``\\ % _`` and matches ``file_path LIKE <folder>% ESCAPE '\\'``:
    a prefix match in which SQLite's LIKE folds ASCII letters only.
    """
    prefixes = []
    for folder in folders:
        f = os.path.abspath(folder)
        if not f.endswith(os.sep):
            f += os.sep
        prefixes.append(f.encode("utf-8").lower())       # bytes.lower() folds ASCII only, as LIKE does
    if lowered is None:
        lowered = [p.encode("utf-8").lower() for p in file_paths]
    pref = tuple(prefixes)
    return np.fromiter((p.startswith(pref) for p in lowered), dtype=bool, count=len(lowered))


def filter_duplicates(results: Result, codes: Dict[str, np.ndarray], tolerance_bits: int = 2) -> Result:
    """Post-filter of image_database.py:1207-1306 given each result's stored sign code.

    ``codes`` maps file_path -> uint8[1152] (absent = no binary embedding: kept).
    Walks results in order; a result within ``tolerance_bits`` differing positions
    of an earlier group's first code is a duplicate: it replaces the group's kept
    entry only if strictly more similar.  The survivors are re-sorted by
    similarity descending (stable).
    """
    group_code: List[np.ndarray] = []     # code that founded each group (never replaced, :1278-1287)
    group_best: List[Tuple[str, float]] = []
    kept: Result = []
    for path, sim in results:
        code = codes.get(path)
        if code is None:
            kept.append((path, sim))
            continue
        hit = -1
        if group_code:
            diff = (np.stack(group_code) != code[None, :]).sum(axis=1)
            near = np.nonzero(diff <= tolerance_bits)[0]
            if near.size:
                hit = int(near[0])
        if hit < 0:
            group_code.append(code)
            group_best.append((path, sim))
            kept.append((path, sim))
        elif sim > group_best[hit][1]:
            old = group_best[hit][0]
            group_best[hit] = (path, sim)
            kept = [(p, s) for p, s in kept if p != old]
            kept.append((path, sim))
    kept.sort(key=lambda r: r[1], reverse=True)
    return kept


class ImageDatabase:
    """Drop-in for the search half of the reference class of the same name."""

    embedding_dim = schema.EMBEDDING_DIM

    def __init__(self, db_path: str, device: int = 0, embedder: Optional[Embedder] = None,
                 nan_policy: str = "reference", verbose: bool = False,
                 binary_score_mode: str = "reference", batch_store: bool = False,
                 devices: Optional[Sequence[int]] = None, hbm_budget_bytes: Optional[int] = None,
                 native_loader: bool = True):
        """``batch_store=True`` also keeps the bf16 copy of the store (half its size again) and sends
        every search — single queries included — through the tensor-core pre-selection + exact
        re-rank: same results, about half the latency per query, and ``search_embeddings`` answers
        many sessions' queries in one pass.  ``devices=[0, 1, ...]`` row-shards the store over
        several GPUs of the box from this one process (``multigpu.MultiGpuIndex``: one launch per
        GPU per query, candidates exchanged over NVLink inside the scan kernel).
        A store whose float32 rows do not fit the GPU next to the bf16 copy is loaded TIERED when
        ``batch_store=True``: the bf16 copy and as many float32 rows as fit stay in HBM, the rest of the float32
        rows go to pinned host memory and are only touched by the exact re-rank (a few hundred rows per query) —
        same answers, about 35M images per B200 become 70M.  ``hbm_budget_bytes`` overrides the free-memory
        reading that decision is based on (tests)."""
        if nan_policy not in ("reference", "exclude"):
            raise ValueError("nan_policy must be 'reference' or 'exclude'")
        if binary_score_mode not in ("reference", "popcount"):
            raise ValueError("binary_score_mode must be 'reference' (uint8 wrap-around, as the reference "
                             "computes it) or 'popcount'")
        import threading
        self._lock = threading.RLock()
        self.binary_score_mode = binary_score_mode
        self.batch_store = bool(batch_store)
        self.hbm_budget_bytes = hbm_budget_bytes
        self.native_loader = bool(native_loader)   # False: always the Python reader (tests compare the two)
        self._placed: List[str] = []
        self.placement = "device"          # or "tiered" (see reload)
        self._codes = None                 # loader.HostCodes once the sign-code fallback is needed
        self._code_mask_key: Optional[Tuple[str, ...]] = None
        self.db_path = db_path
        self.embedder = embedder
        self.nan_policy = nan_policy
        self.verbose = verbose
        if devices is not None and len(devices) > 1:
            from .multigpu import MultiGpuIndex
            self.index = MultiGpuIndex(devices)
        else:
            self.index = GpuIndex(device if not devices else int(devices[0]))
        self._paths: List[str] = []
        self._lowered: Optional[List[bytes]] = None
        self._image_ids = np.zeros(0, dtype=np.int64)
        self._rowids = np.zeros(0, dtype=np.int64)
        self._mtimes = np.zeros(0, dtype=np.float64)
        self._alive: Optional[np.ndarray] = None      # False = the row's mapping is gone (refresh): masked out
        self._watch = None
        self._data_version = -1
        self._dangling = 0            # mapped rowids (<= the last resident one) that have no vec0 row: see refresh()
        self.load_seconds = 0.0
        self.load_source = ""
        self.reloads = 0
        self._binary_count = 0
        self._vec0_count = 0
        self._mask_key: Optional[Tuple[str, ...]] = None
        self.reload()

    # ---- store ----------------------------------------------------------------------
    def _log(self, *a) -> None:
        if self.verbose:
            print(*a, flush=True)

    def _pos_of(self, rowids) -> np.ndarray:
        """Scan positions of resident rowids (``self._rowids`` is ascending: the scan order)."""
        return np.searchsorted(self._rowids, np.asarray(rowids, dtype=np.int64))

    def _load_range(self, conn, index, lo=None, hi=None, rows_hint: Optional[int] = None) -> loader.HostStore:
        """Append the joined vec0 rows with rowid in (lo, hi] to ``index`` (a ``GpuIndex``) and return their
        metadata (``rows=None``).  ``rows_hint`` = reserve a fresh store for about that many rows first (a load);
        None = append to the store that is there (a refresh).  Host memory stays O(chunk) either way:
          * native reader (``clipdb_append_sqlite``) when vec0 is a plain table: SQLite's C library copies every
            blob once, into a pinned double buffer, and the DMA of a chunk overlaps the reading of the next;
          * otherwise (sqlite-vec's virtual table or its shadow tables) the Python reader: ``fetchmany`` chunks
            copied into the context's pinned staging buffer."""
        state = {"reserved": rows_hint is None}

        def reserve(dim: int, at_least: int) -> None:
            if not state["reserved"]:
                self._placed.append(self._reserve(index, max(rows_hint, at_least, 1), dim))
                state["reserved"] = True

        if self.native_loader and loader.vec0_source(conn) == "plain-table":
            dim = loader.peek_dim(conn, lo, hi)
            if dim is not None:
                reserve(dim, 1)
                ids, images, mtimes, paths = [], [], [], []

                def on_chunk(c_ids, c_images, c_mtimes, c_paths):
                    ids.append(c_ids)
                    images.append(c_images)
                    mtimes.append(c_mtimes)
                    paths.extend(c_paths)
                try:
                    vec0_rows, joined = index.append_sqlite(self.db_path, lo, hi, on_chunk, loader.CHUNK_ROWS)
                except _lib.ClipdbError as e:
                    if e.code != _lib.ERR_UNSUPPORTED:
                        raise
                else:
                    cat = lambda parts, dt: np.concatenate(parts) if parts else np.zeros(0, dtype=dt)   # noqa: E731
                    return loader.HostStore(cat(ids, np.int64), None, cat(images, np.int64), paths,
                                            loader.count_binary(conn), vec0_rows, "plain-table (native reader)",
                                            vec0_rows - joined, cat(mtimes, np.float64), dim)
        stage = {"buf": None}

        def sink(chunk: loader.StoreChunk) -> None:
            m, dim = len(chunk), chunk.dim
            reserve(dim, m)
            if stage["buf"] is None or stage["buf"].shape[1] != dim or stage["buf"].shape[0] < m:
                stage["buf"] = index.stage_buffer(max(loader.CHUNK_ROWS, m), dim)
            buf = stage["buf"][:m]
            chunk.write_rows(buf)                      # blobs -> pinned memory, one memcpy per row
            index.append(buf, chunk.rowids)
            index.synchronize()                        # the staging buffer is reused by the next chunk
        return loader.stream_store(self.db_path, sink, expect_dim=(index.dim or None) if rows_hint is None else None,
                                   min_rowid=lo, max_rowid=hi, conn=conn)

    def _reserve(self, index, rows_hint: int, dim: int) -> str:
        """Reserve room for ``rows_hint`` rows on ``index``: all in HBM when they fit (next to the bf16 copy if one
        is wanted), else — with ``batch_store`` — tiered.  Returns "device" or "tiered"."""
        need32 = rows_hint * dim * 4
        need16 = rows_hint * dim * 2 if (self.batch_store and dim == schema.EMBEDDING_DIM) else 0
        free = self.hbm_budget_bytes if self.hbm_budget_bytes is not None else index.get_option("device_free_bytes")
        slack = min(2 << 30, free // 8)            # workspaces, candidate lists, fragmentation
        if need32 + need16 + slack <= free:
            index.reserve(rows_hint, dim, explicit_rowids=True)
            return "device"
        if need16 and need16 + slack < free:
            device_rows = max(0, (free - need16 - slack) // (dim * 4))
            index.reserve(rows_hint, dim, explicit_rowids=True, placement="host", device_rows=device_rows)
            index.enable_batch()                   # every append converts its rows: one pass builds both copies
            index.set_option("batch_min_nq", 1)    # every search pre-selects on the resident bf16 copy
            return "tiered"
        raise MemoryError(
            f"{rows_hint} rows x {dim} float32 = {need32 / 1e9:.1f} GB do not fit the GPU ({free / 1e9:.1f} GB free)"
            + ("" if need16 else "; batch_store=True keeps only a bf16 copy resident (half the bytes) and tiers the "
                                 "float32 rows into host memory") + "; or shard over more GPUs with devices=[...]")

    @_serialised
    def reload(self) -> None:
        """(Re)read the whole database into HBM, streamed: every chunk of rows goes from SQLite through a pinned
        staging buffer straight into the resident store.  With several GPUs the shard boundaries are computed
        once and every range is read through the same connection inside ONE read transaction, so a scan writing
        concurrently cannot make two shards disagree about where one ends and the next begins."""
        import time
        t0 = time.perf_counter()
        multi = hasattr(self.index, "shards")
        world = self.index.world if multi else 1
        with loader.snapshot(self.db_path) as conn:
            self._data_version = loader.data_version(self._watch_conn())
            mapped = int(conn.execute("SELECT COUNT(*) FROM image_embeddings").fetchone()[0])
            parts = []
            self._placed = []
            if multi and mapped > 0:
                ranges, n_mapped = loader.plan_shards(conn, world)
                if 0 < n_mapped < world:
                    raise ValueError(f"{n_mapped} rows cannot be sharded over {world} GPUs")
                for rank, (lo, hi) in enumerate(ranges):
                    host = self._load_range(conn, self.index.shards[rank], lo, hi, rows_hint=n_mapped // world + 1)
                    if host.rowids.shape[0] < 1:
                        raise ValueError("every shard needs at least one row (too many vec0 rows are missing)")
                    self.index.note_shard_loaded(rank, host.rowids.shape[0])
                    parts.append(host)
                self.index.finish_load()
            else:
                single = self.index.shards[0] if multi else self.index
                parts.append(self._load_range(conn, single, rows_hint=mapped))
            self._vec0_count = loader.count_vec0(conn)
        first = parts[0]
        self.load_source = first.source
        self._binary_count = first.binary_count
        self._paths = [fp for p in parts for fp in p.file_paths]
        self._lowered = None
        self._image_ids = np.concatenate([p.image_ids for p in parts])
        self._mtimes = np.concatenate([p.mtimes for p in parts])
        self._rowids = np.concatenate([p.rowids for p in parts])
        self._alive = None
        self._mask_key = None
        self._codes = None
        self._code_mask_key = None
        n = self._rowids.shape[0]
        self._dangling = 0
        self.placement = "tiered" if "tiered" in self._placed else "device"
        if self.batch_store and n and self.index.dim == schema.EMBEDDING_DIM:
            self.index.enable_batch()
            if not multi:
                self.index.set_option("batch_min_nq", 1)
            else:
                self.index.prefer_batch = self.placement == "tiered"
        self.load_seconds = time.perf_counter() - t0
        self.reloads += 1
        if loader.data_version(self._watch_conn()) != self._data_version:
            # someone committed while the load ran (the native reader and this connection hold separate snapshots):
            # reconcile with what a fresh connection sees now
            self.refresh()
        self._log(f"loaded {n} rows ({first.source}, {self.placement} store) in {self.load_seconds:.2f} s "
                  f"({n / max(self.load_seconds, 1e-9):.0f} rows/s); "
                  f"{sum(p.dropped for p in parts)} vec0 rows without a mapping were skipped")

    def _match_mapping(self, mp: "loader.Mapping"):
        """(index of every resident rowid in the mapping, which resident rowids are still mapped, how many mapped
        rowids up to the last resident one are NOT resident)."""
        old = self._rowids
        n_old = old.shape[0]
        if n_old == 0 or len(mp.rowids) == 0:
            return np.zeros(n_old, dtype=np.int64), np.zeros(n_old, dtype=bool), 0
        at = np.searchsorted(mp.rowids, old)
        at[at >= len(mp.rowids)] = len(mp.rowids) - 1
        present = mp.rowids[at] == old
        mapped_up_to_last = int(np.searchsorted(mp.rowids, int(old[-1]), side="right"))
        return at, present, mapped_up_to_last - int(present.sum())

    def _watch_conn(self):
        """A connection kept open only to read ``PRAGMA data_version``: it changes iff someone else committed."""
        if self._watch is None:
            self._watch = loader.connect(self.db_path)
        return self._watch

    @_serialised
    def refresh(self, force: bool = False) -> int:
        """Bring the resident store in line with what a fresh connection would see now — what the reference does
        implicitly by reopening SQLite for every search (image_database.py:1475).  Free when nothing was
        committed since the last look (``PRAGMA data_version``).  Otherwise one pass over the integer mapping
        (``image_embeddings JOIN images``: no blobs) finds
          * rows the scanner appended (new vec0 rowids are always larger: ``INSERT INTO vec0`` auto-assigns,
            :1171-1175)                                           -> streamed in and appended;
          * rows whose mapping disappeared — a modified file is re-keyed by ``INSERT OR REPLACE INTO images``
            (:1137-1148), which orphans its old vec0 row; the reference's INNER JOINs then drop it
            (:1569-1570)                                          -> masked out of every search;
          * rows whose image row changed (``last_modified`` / ``image_id``), the trace of an in-place
            ``UPDATE vec0`` (:1165-1167)                            -> blob re-read, ``clipdb_update_row``;
          * anything else (a mapping that appeared for an old rowid)  -> full ``reload()``.
        Returns the number of rows appended + updated + retired."""
        dv = loader.data_version(self._watch_conn())
        if not force and dv == self._data_version:
            return 0
        multi = hasattr(self.index, "shards")
        touched = 0
        with loader.snapshot(self.db_path) as conn:
            self._data_version = dv
            mp = loader.read_mapping(conn)
            old = self._rowids
            n_old = old.shape[0]
            last = int(old[-1]) if n_old else None
            at, present, dangling = self._match_mapping(mp)
            n_old_in_map = int(np.searchsorted(mp.rowids, last, side="right")) if last is not None else 0
            if n_old == 0 or dangling != self._dangling:
                # a store that was empty, or a mapping that appeared for a rowid we skipped at load time.  Mapped
                # rowids that still have no vec0 row after a fresh load are permanent: remembered, so that they do
                # not trigger a reload on every later change of the database
                conn.execute("ROLLBACK")
                self.reload()
                self._dangling = self._match_mapping(mp)[2]
                return self._rowids.shape[0]
            # retired rows (and rows whose mapping came back)
            alive_before = self._alive if self._alive is not None else np.ones(n_old, dtype=bool)
            flipped = alive_before != present
            if flipped.any():
                self._alive = None if present.all() else present.copy()
                self._mask_key = None
                touched += int(flipped.sum())
            # in-place re-embeddings
            changed = present & ((mp.image_ids[at] != self._image_ids) | (mp.mtimes[at] != self._mtimes))
            if changed.any():
                for chunk in loader.read_rows_by_rowid(conn, old[changed].tolist()):
                    pos = self._pos_of(chunk.rowids)
                    for j, p in enumerate(pos.tolist()):
                        self.index.update_row(p, chunk.rows[j])
                        self._paths[p] = chunk.file_paths[j]
                    self._image_ids[pos] = chunk.image_ids
                    self._mtimes[pos] = chunk.mtimes
                    touched += len(pos)
                self._lowered = None
                self._mask_key = None
            # appended rows
            if len(mp.rowids) > n_old_in_map:
                target = self.index.shards[-1] if multi else self.index
                host = self._load_range(conn, target, lo=last)
                m = host.rowids.shape[0]
                if m:
                    if multi:
                        self.index.note_appended(m)
                    self._paths.extend(host.file_paths)
                    self._lowered = None
                    self._image_ids = np.concatenate([self._image_ids, host.image_ids])
                    self._mtimes = np.concatenate([self._mtimes, host.mtimes])
                    self._rowids = np.concatenate([self._rowids, host.rowids])
                    if self._alive is not None:
                        self._alive = np.concatenate([self._alive, np.ones(m, dtype=bool)])
                    self._mask_key = None
                    touched += m
            binary_count = loader.count_binary(conn)
            if binary_count != self._binary_count or touched:
                self._codes = None            # the sign-code fallback store is re-read on next use
                self._code_mask_key = None
            self._binary_count = binary_count
            self._vec0_count = loader.count_vec0(conn)
        return touched

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self) -> None:
        if self._watch is not None:
            self._watch.close()
            self._watch = None
        self.index.close()

    # ---- search -----------------------------------------------------------------------
    def _embed(self, query: str, is_image: bool, what: str) -> Optional[np.ndarray]:
        if isinstance(query, str) and query.lower().startswith(VECTOR_PREFIX):
            return load_vector(query[len(VECTOR_PREFIX):].strip(), self.embedding_dim)
        if self.embedder is None:
            raise RuntimeError("no embedder configured: pass embedder=... or call search_embedding() "
                               "with float32 vectors (the SigLIP model is outside this package)")
        return self.embedder.image(query) if is_image else self.embedder.text(query)

    def search(self, query: str, k: int = 10, is_image_path: bool = False,
               query2: str = None, is_image_path2: bool = False,
               weights: Tuple[float, float] = (0.5, 0.5),
               negative_query: str = None, negative_is_image: bool = False,
               negative_weight: float = 0.5,
               negative_queries: List[str] = None, negative_is_images: List[bool] = None,
               negative_weights: List[float] = None,
               filter_folders: List[str] = None,
               profile: bool = False,
               show_duplicates: bool = False) -> Result:
        """Same contract as image_database.py:1308-1337."""
        # first / second positive query (:1340-1376)
        if is_image_path and not os.path.exists(query):
            print(f"Error: Image file {query} does not exist")
            return []
        e1 = self._embed(query, is_image_path, "query")
        if e1 is None:
            print("Error: Failed to generate embedding from image")
            return []
        e2 = None
        if query2 is not None:
            if is_image_path2 and not os.path.exists(query2):
                print(f"Error: Image file {query2} does not exist")
                return []
            e2 = self._embed(query2, is_image_path2, "second query")
            if e2 is None:
                print("Error: Failed to generate embedding from second image")
                return []
        # negatives: legacy single first, then the list; missing image files are skipped (:1402-1451)
        negs, neg_ws = [], []
        if negative_query is not None:
            if negative_is_image and not os.path.exists(negative_query):
                print(f"Warning: Negative image file {negative_query} does not exist, ignoring negative prompt")
            else:
                v = self._embed(negative_query, negative_is_image, "negative")
                if v is not None:
                    negs.append(v)
                    neg_ws.append(negative_weight)
        if negative_queries is not None:
            for i, nq in enumerate(negative_queries):
                is_img = negative_is_images[i] if negative_is_images and i < len(negative_is_images) else False
                w = negative_weights[i] if negative_weights and i < len(negative_weights) else negative_weight
                if is_img and not os.path.exists(nq):
                    print(f"Warning: Negative image file {nq} does not exist, skipping")
                    continue
                v = self._embed(nq, is_img, "negative")
                if v is not None:
                    negs.append(v)
                    neg_ws.append(w)
        return self.search_embedding(e1, k=k, embedding2=e2, weights=weights, negative_embeddings=negs,
                                     negative_weights=neg_ws, filter_folders=filter_folders,
                                     profile=profile, show_duplicates=show_duplicates)

    @_serialised
    def search_embeddings(self, embeddings, k: int = 10, filter_folders: Optional[Sequence[str]] = None
                          ) -> List[Result]:
        """Many ready-made query embeddings (e.g. one per interactive session) in one call: with
        ``batch_store=True`` they share one pass over the store.  Per query the same list as
        ``search_embedding(..., show_duplicates=True)``: (file_path, similarity), best first."""
        q = np.ascontiguousarray(embeddings, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.embedding_dim:
            raise ValueError(f"embeddings must be [nq, {self.embedding_dim}]")
        if self._binary_count <= 0 or self._vec0_count <= 0 or self.index.num_rows == 0:
            return [self.search_embedding(v, k=k, filter_folders=filter_folders, show_duplicates=True) for v in q]
        k = int(k)
        if k < 0:
            k = self.index.num_rows
        use_mask = self._install_mask(filter_folders)
        multi = hasattr(self.index, "shards")
        if multi and getattr(self.index, "batch_enabled", False) and not use_mask and \
                1 <= k <= min(128, min(hi - lo for lo, hi in self.index.bounds)):
            res = self.index.search_batch(q, k)          # one batched pass per 256 queries on every GPU
            multi = False
        else:
            res = None if multi else self.index.search(q, k, use_mask=use_mask)
        out: List[Result] = []
        for i in range(q.shape[0]):
            if multi:
                rowids, dist, nan = self.index.search_any_k(q[i], k, use_mask)
            else:
                (rowids, dist), nan = res.row(i), int(res.nan_rows[i])
            if nan > 0 and self.nan_policy == "reference" and k > 0:
                out.append([])            # the reference's search() returns [] (see search_embedding)
                continue
            out.append([(self._paths[p], 1.0 - float(d)) for p, d in zip(self._pos_of(rowids).tolist(), dist)])
        return out

    def _install_mask(self, filter_folders: Optional[Sequence[str]]) -> bool:
        """Admission bitset = the folder WHERE clause AND the rows whose mapping still exists."""
        retired = self._alive is not None and not self._alive.all()
        if not filter_folders and not retired:
            return False
        key = (tuple(filter_folders or ()), retired)
        if key != self._mask_key:
            admitted = np.ones(len(self._paths), dtype=bool)
            if filter_folders:
                if self._lowered is None:
                    self._lowered = [p.encode("utf-8").lower() for p in self._paths]
                admitted = like_prefix_mask(self._paths, filter_folders, self._lowered)
            if retired:
                admitted = admitted & self._alive
            self.index.set_mask(admitted)
            self._mask_key = key
        return True

    @_serialised
    def search_embedding(self, embedding1, k: int = 10, embedding2=None,
                         weights: Tuple[float, float] = (0.5, 0.5),
                         negative_embeddings: Sequence = (), negative_weights: Sequence[float] = (),
                         filter_folders: Optional[Sequence[str]] = None, profile: bool = False,
                         show_duplicates: bool = False) -> Result:
        """``search()`` after the embedding step: blend, negatives, guards, scan, top-k,
        similarity conversion, duplicate filter (image_database.py:1378-1658)."""
        import time
        timings = {}
        # guards (:1488-1500, :1532-1555)
        if self._binary_count <= 0:
            print("Error: Database has no embeddings. Please run scan first.")
            return []
        if self._vec0_count <= 0:
            # vec0 is empty: the sign-code fallback (:1591-1629)
            results, codes = self._binary_fallback(embedding1, k, embedding2, weights, negative_embeddings,
                                                   negative_weights, filter_folders, timings)
            return self._finish(results, show_duplicates, profile, timings, codes=codes)
        try:
            k = int(k)
            t0 = time.time()
            use_mask = self._install_mask(filter_folders)
            timings["build_query"] = time.time() - t0
            t0 = time.time()
            if k < 0:
                k = self.index.num_rows          # SQLite: negative LIMIT = no limit
            if self.index.num_rows == 0:
                return []
            res = self.index.blend_search(embedding1, k, e2=embedding2, weights=weights,
                                          negatives=negative_embeddings,
                                          negative_weights=negative_weights, use_mask=use_mask)
            rowids, dist = res.row(0)
            if res.nan_rows[0] > 0 and self.nan_policy == "reference" and k > 0:
                # SQLite stores a NaN distance as NULL, NULLs sort first, and the reference's
                # `1.0 - distance` then raises inside its try block (:1588, :1637-1640)
                raise TypeError("unsupported operand type(s) for -: 'float' and 'NoneType'")
            positions = self._pos_of(rowids).tolist()
            top = [(self._paths[p], 1.0 - float(d)) for p, d in zip(positions, dist)]
            timings["db_query"] = time.time() - t0
            results = [(p, float(s)) for p, s in top]
            image_ids = {self._paths[p]: int(self._image_ids[p]) for p in positions}
        except Exception as e:                       # error envelope, :1637-1640
            print(f"Error during search: {e}")
            return []
        return self._finish(results, show_duplicates, profile, timings, image_ids=image_ids)

    # ---- sign-code fallback (vec0 empty, image_database.py:1591-1629) -------------------------
    def _filtered_statement_walks_the_path_index(self) -> bool:
        """Ask THIS SQLite how it would run the fallback's statement with a folder WHERE clause.
        3.45 walks the UNIQUE index on images.file_path (rows arrive in file_path order) instead
        of scanning binary_embeddings (rowid order); the arrival order is the tie-break of the
        reference's stable sort, so it has to be mirrored."""
        conn = loader.connect(self.db_path)
        try:
            plan = conn.execute("EXPLAIN QUERY PLAN " + loader.BINARY_SQL +
                                " WHERE (i.file_path LIKE ? ESCAPE '\\')", ("x%",)).fetchall()
        finally:
            conn.close()
        # "SCAN i USING [COVERING] INDEX sqlite_autoindex_images_1" = file_path order; a bare "SCAN i" would be a
        # table scan in id order, "SCAN be" the unfiltered statement's binary_embeddings rowid order
        first = str(plan[0][-1]) if plan else ""
        return first.startswith("SCAN i") and "INDEX" in first and "autoindex_images" in first

    def _binary_fallback(self, embedding1, k, embedding2, weights, negative_embeddings, negative_weights,
                         filter_folders, timings) -> Result:
        import time
        try:
            t0 = time.time()
            if self._codes is None:
                self._codes = loader.read_codes(self.db_path, expect_dim=self.embedding_dim)
                self.index.load_codes(self._codes.codes)
                self._code_mask_key = None
            paths = self._codes.file_paths
            n = len(paths)
            use_mask = False
            admitted_n = n
            if filter_folders:
                key = tuple(filter_folders)
                if key != self._code_mask_key:
                    admitted = like_prefix_mask(paths, filter_folders)
                    order = None
                    if self._filtered_statement_walks_the_path_index():
                        by_path = sorted(range(n), key=lambda i: paths[i].encode("utf-8"))   # BINARY collation
                        order = np.empty(n, dtype=np.uint32)
                        order[by_path] = np.arange(n, dtype=np.uint32)
                    self.index.set_code_mask(admitted, order)
                    self._code_mask_key = key
                    self._code_admitted = int(admitted.sum())
                use_mask = True
                admitted_n = self._code_admitted
            timings["build_query"] = time.time() - t0
            t0 = time.time()
            # blend / negatives exactly as for the float path (:1378-1472), then the sign code (:1593)
            query, _flags = self.index.blend(embedding1, embedding2, weights, negative_embeddings, negative_weights)
            code = (query >= 0).astype(np.uint8)
            k = int(k)
            kk = k if k >= 0 else max(admitted_n + k, 0)      # `candidate_scores[:k]` is a Python slice (:1628)
            pos, scores = self.index.binary_search(code, kk, score_mode=self.binary_score_mode, use_mask=use_mask)
            results = [(paths[int(p)], float(int(s)) / self.embedding_dim) for p, s in zip(pos, scores)]
            timings["db_query"] = time.time() - t0
            # the stored sign codes of the results are already in host memory (the duplicate filter needs them)
            return results, {paths[int(p)]: self._codes.codes[int(p)] for p in pos}
        except Exception as e:                       # error envelope, :1637-1640
            print(f"Error during search: {e}")
            return [], {}

    def _finish(self, results: Result, show_duplicates: bool, profile: bool, timings,
                image_ids: Optional[Dict[str, int]] = None, codes: Optional[Dict[str, np.ndarray]] = None) -> Result:
        import time
        if not show_duplicates and len(results) > 0:
            t0 = time.time()
            results = self._filter_duplicates(results, tolerance_bits=2, image_ids=image_ids, codes=codes)
            timings["filter_duplicates"] = time.time() - t0
        if profile and timings:
            print("\n=== Search Performance Profile ===")
            total = sum(timings.values())
            for op, dur in sorted(timings.items(), key=lambda x: x[1], reverse=True):
                pct = (dur / total * 100) if total > 0 else 0
                print(f"  {op:25s}: {dur * 1000:7.2f}ms ({pct:5.1f}%)")
            print(f"  {'TOTAL':25s}: {total * 1000:7.2f}ms")
            print("=" * 40 + "\n")
        return results

    def _filter_duplicates(self, results: Result, tolerance_bits: int = 2,
                           image_ids: Optional[Dict[str, int]] = None,
                           codes: Optional[Dict[str, np.ndarray]] = None) -> Result:
        """Apply ``filter_duplicates`` to the k results' stored sign codes.  The reference fetches
        them per search with one ``SELECT id FROM images WHERE file_path = ?`` per result and one
        ``IN (...)`` query (image_database.py:1232-1255); here the image ids come from the join
        made at load time (same ids: ``ie.image_id = i.id``), so only the ``IN`` query remains, and
        the sign-code fallback passes the codes it already holds."""
        if codes is not None:
            before = len(results)
            out = filter_duplicates(results, codes, tolerance_bits)
            if len(out) < before:
                print(f"Filtered out {before - len(out)} duplicate(s) (tolerance: {tolerance_bits} bits)")
            return out
        conn = loader.connect(self.db_path)
        try:
            ids = dict(image_ids) if image_ids is not None else {}
            if image_ids is None:
                for path, _ in results:
                    row = conn.execute("SELECT id FROM images WHERE file_path = ?", (path,)).fetchone()
                    if row:
                        ids[path] = row[0]
            codes = {}
            if ids:
                id_list = list(ids.values())
                by_id = {}
                for lo in range(0, len(id_list), 500):      # stay under SQLite's bound-variable limit
                    part = id_list[lo:lo + 500]
                    marks = ",".join("?" * len(part))
                    for image_id, blob in conn.execute(
                            f"SELECT image_id, embedding FROM binary_embeddings WHERE image_id IN ({marks})", part):
                        by_id[image_id] = np.frombuffer(blob, dtype=np.uint8)
                codes = {p: by_id[i] for p, i in ids.items() if i in by_id}
        finally:
            conn.close()
        before = len(results)
        out = filter_duplicates(results, codes, tolerance_bits)
        if len(out) < before:
            print(f"Filtered out {before - len(out)} duplicate(s) (tolerance: {tolerance_bits} bits)")
        return out
