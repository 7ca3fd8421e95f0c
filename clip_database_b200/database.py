"""``ImageDatabase`` — the reference's search surface on top of the GPU index.

Mirrors ``ImageDatabase.search`` (image_database.py:1308-1658): same signature,
same result type (``List[Tuple[file_path, similarity]]`` sorted by similarity
descending), same guards and fallbacks — with the body replaced: the store is
loaded once from the SQLite database into resident HBM and every search is one
blend + scan + top-k on the GPU instead of a per-row SQL function call.

The SigLIP model is out of scope (no weights offline): ``search()`` takes text /
image-path queries only when an ``embedder`` is supplied; ``search_embedding()``
takes the float32[1152] vectors directly and is what ``search()`` calls after
embedding.  There is no CPU search path.
"""
from __future__ import annotations

import os
import sqlite3
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import loader, schema
from .index import GpuIndex

Result = List[Tuple[str, float]]


class Embedder:
    """Interface for the out-of-scope model: text / image path -> float32[dim] (or None)."""

    def text(self, query: str) -> Optional[np.ndarray]:       # image_database.py:509-543
        raise NotImplementedError

    def image(self, path: str) -> Optional[np.ndarray]:       # image_database.py:443-463
        raise NotImplementedError


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
                 devices: Optional[Sequence[int]] = None):
        """``batch_store=True`` also keeps the bf16 copy of the store (half its size again) and sends
        every search — single queries included — through the tensor-core pre-selection + exact
        re-rank: same results, about half the latency per query, and ``search_embeddings`` answers
        many sessions' queries in one pass.  ``devices=[0, 1, ...]`` row-shards the store over
        several GPUs of the box from this one process (``multigpu.MultiGpuIndex``: one launch per
        GPU per query, candidates exchanged over NVLink inside the scan kernel)."""
        if nan_policy not in ("reference", "exclude"):
            raise ValueError("nan_policy must be 'reference' or 'exclude'")
        if binary_score_mode not in ("reference", "popcount"):
            raise ValueError("binary_score_mode must be 'reference' (uint8 wrap-around, as the reference "
                             "computes it) or 'popcount'")
        self.binary_score_mode = binary_score_mode
        self.batch_store = bool(batch_store)
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
        self._rowid_to_pos: Dict[int, int] = {}
        self._binary_count = 0
        self._vec0_count = 0
        self._mask_key: Optional[Tuple[str, ...]] = None
        self.reload()

    # ---- store ----------------------------------------------------------------------
    def _log(self, *a) -> None:
        if self.verbose:
            print(*a, flush=True)

    def reload(self) -> None:
        """(Re)read the whole database into HBM.  With several GPUs every shard's rowid range is read
        and uploaded on its own, so the float32 matrix never has to fit host memory at once."""
        multi = hasattr(self.index, "shards")
        world = self.index.world if multi else 1
        sharded_load = multi
        if multi:
            n_joined = loader.shard_rowid_range(self.db_path, 0, world)[2]
            if 0 < n_joined < world:
                raise ValueError(f"{n_joined} rows cannot be sharded over {world} GPUs")
            sharded_load = n_joined > 0              # a binary-only database has no float rows to shard
        parts = []
        for rank in range(world if sharded_load else 1):
            if sharded_load:
                lo, hi, _ = loader.shard_rowid_range(self.db_path, rank, world)
                host = loader.read_store(self.db_path, min_rowid=lo, max_rowid=hi)
                self.index.load_shard(rank, host.rows, host.rowids)
                host.rows = None                     # uploaded: free the host copy before the next range
            else:
                host = loader.read_store(self.db_path)
            parts.append(host)
        if sharded_load:
            self.index.finish_load()
        first = parts[0]
        self._binary_count = first.binary_count
        self._vec0_count = sum(p.vec0_count for p in parts)
        self._paths = [fp for p in parts for fp in p.file_paths]
        self._lowered = None
        self._image_ids = np.concatenate([p.image_ids for p in parts])
        self._rowids = np.concatenate([p.rowids for p in parts])
        self._rowid_to_pos = {int(r): i for i, r in enumerate(self._rowids)}
        self._mask_key = None
        self._codes = None
        self._code_mask_key = None
        if not multi and first.rows.shape[0]:
            self.index.load(first.rows, first.rowids)
        if self.batch_store and self._rowids.shape[0] and self.index.dim == schema.EMBEDDING_DIM:
            self.index.enable_batch()
            if not multi:
                self.index.set_option("batch_min_nq", 1)
        self._log(f"loaded {self._rowids.shape[0]} rows ({first.source}); "
                  f"{sum(p.dropped for p in parts)} vec0 rows without a mapping were skipped")

    def refresh(self) -> int:
        """Append rows the scanner added since the last load (new vec0 rowids are always
        larger: INSERT INTO vec0 auto-assigns, image_database.py:1171-1175).  Returns
        the number of rows appended.  In-place UPDATEs of old rows need ``reload()``."""
        last = int(self._rowids[-1]) if len(self._paths) else None
        host = loader.read_store(self.db_path, expect_dim=self.index.dim or None, min_rowid=last)
        if host.binary_count != self._binary_count:
            self._codes = None            # sign codes were added: the fallback store is re-read on next use
            self._code_mask_key = None
        self._binary_count = host.binary_count
        self._vec0_count = self._vec0_count + host.vec0_count if last is not None else host.vec0_count
        m = host.rows.shape[0]
        if m == 0:
            return 0
        if not self._paths:
            self.index.load(host.rows, host.rowids)
        else:
            self.index.append(host.rows, host.rowids)
        base = len(self._paths)
        self._paths.extend(host.file_paths)
        self._lowered = None
        self._image_ids = np.concatenate([self._image_ids, host.image_ids])
        self._rowids = np.concatenate([self._rowids, host.rowids]) if base else host.rowids
        for i, r in enumerate(host.rowids):
            self._rowid_to_pos[int(r)] = base + i
        self._mask_key = None
        return m

    def close(self) -> None:
        self.index.close()

    # ---- search -----------------------------------------------------------------------
    def _embed(self, query: str, is_image: bool, what: str) -> Optional[np.ndarray]:
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
        if multi and getattr(self.index, "batch_enabled", False) and not use_mask and 1 <= k <= 128:
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
            out.append([(self._paths[self._rowid_to_pos[int(r)]], 1.0 - float(d)) for r, d in zip(rowids, dist)])
        return out

    def _install_mask(self, filter_folders: Optional[Sequence[str]]) -> bool:
        if not filter_folders:
            return False
        key = tuple(filter_folders)
        if key != self._mask_key:
            if self._lowered is None:
                self._lowered = [p.encode("utf-8").lower() for p in self._paths]
            self.index.set_mask(like_prefix_mask(self._paths, filter_folders, self._lowered))
            self._mask_key = key
        return True

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
            positions = [self._rowid_to_pos[int(r)] for r in rowids]
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
        return bool(plan) and str(plan[0][-1]).startswith("SCAN i")

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
