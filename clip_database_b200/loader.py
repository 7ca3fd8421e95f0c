"""SQLite -> host arrays: read a reference-schema database for loading into HBM.

Reproduces the row set and order of the reference's search statement
(image_database.py:1564-1571): a full scan of ``vec0`` in rowid order,
INNER JOINed to ``image_embeddings`` (``vec0.rowid = ie.rowid``) and ``images``
(``ie.image_id = i.id``).  vec0 rows without both partners are dropped here, as
SQLite drops them there.

Where the float32 vectors come from, in order of preference:
  1. the real virtual table through sqlite-vec, when ``import sqlite_vec`` works;
  2. sqlite-vec's shadow tables read directly (``vec0_rowids`` / ``vec0_chunks`` /
     ``vec0_vector_chunks00``) [UPSTREAM-UNVERIFIED layout, see schema.py];
  3. a plain table named ``vec0(rowid, embedding BLOB)`` (synthetic stand-in).

Everything streams: ``iter_store`` yields chunks of a few thousand rows (one ``fetchmany`` of the vec0
cursor, one rowid-range query of the mapping, numpy for the join), so a 10M-row database is never held
in host memory — ``ImageDatabase`` copies each chunk into a pinned staging buffer and appends it to
the resident store.  All statements of one load run on ONE connection inside ONE read transaction
(``snapshot``), so the rows, the mapping and the shard boundaries of a multi-GPU load describe the same
database state even while a scan is writing (the reference runs SQLite in WAL mode).
"""
from __future__ import annotations

import sqlite3
from contextlib import contextmanager
from dataclasses import dataclass, field
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

CHUNK_ROWS = 8192
_MIN_ROWID = -(1 << 63)
_MAX_ROWID = (1 << 63) - 1


class StoreChunk:
    """Consecutive joined rows in scan order.  The vectors are held as the blobs SQLite returned; ``write_rows``
    copies them straight into a caller buffer (the pinned staging buffer of the GPU load: no intermediate joined
    copy), ``rows`` materialises a float32 matrix for host-side users."""
    __slots__ = ("rowids", "image_ids", "mtimes", "file_paths", "blobs", "dim", "_rows")

    def __init__(self, rowids, image_ids, mtimes, file_paths, blobs=None, dim=0, rows=None):
        self.rowids = rowids          # int64 [m], ascending
        self.image_ids = image_ids    # int64 [m]
        self.mtimes = mtimes          # float64 [m]  images.last_modified (refresh compares it)
        self.file_paths = file_paths  # [m]
        self.blobs = blobs            # m little-endian float32 blobs of dim * 4 bytes (or None when built from rows)
        self.dim = int(dim if rows is None else rows.shape[1])
        self._rows = rows

    def __len__(self) -> int:
        return len(self.rowids)

    @property
    def rows(self) -> np.ndarray:
        """float32 [m, dim] (read-only view of one joined copy of the blobs)."""
        if self._rows is None:
            self._rows = np.frombuffer(b"".join(self.blobs), dtype="<f4").reshape(len(self.blobs), self.dim)
        return self._rows

    def write_rows(self, out: np.ndarray) -> None:
        """Copy the vectors into ``out`` (C-contiguous float32 ``[m, dim]``), one memcpy per row."""
        m = len(self.rowids)
        if out.shape != (m, self.dim) or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 [m, dim] array")
        if self._rows is not None or self.blobs is None:
            out[:] = self.rows
            return
        dst = memoryview(out).cast("B")
        step = self.dim * 4
        at = 0
        for blob in self.blobs:
            dst[at:at + step] = blob
            at += step


@dataclass
class StoreStats:
    """Filled while ``iter_store`` runs."""
    source: str = ""
    vec0_rows: int = 0        # vec0 rows seen in the range, before the joins
    joined_rows: int = 0
    dim: Optional[int] = None


@dataclass
class HostStore:
    rowids: np.ndarray        # int64 [n], ascending (= scan order)
    rows: Optional[np.ndarray]  # float32 [n, dim] (None when the rows were streamed to a sink)
    image_ids: np.ndarray     # int64 [n]
    file_paths: List[str]     # [n]
    binary_count: int         # COUNT(*) FROM binary_embeddings (guard, image_database.py:1488-1500)
    vec0_count: int           # COUNT(*) FROM vec0 before the joins (guard, :1532-1540)
    source: str               # which of the three vec0 readers was used
    dropped: int              # vec0 rows excluded by the INNER JOINs
    mtimes: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.float64))
    dim_: int = 0

    @property
    def dim(self) -> int:
        if self.rows is not None and self.rows.ndim == 2:
            return int(self.rows.shape[1])
        return self.dim_


@dataclass
class Mapping:
    """``image_embeddings JOIN images`` without blobs or paths: what ``refresh`` reconciles against."""
    rowids: np.ndarray        # int64, ascending
    image_ids: np.ndarray     # int64
    mtimes: np.ndarray        # float64


@dataclass
class HostCodes:
    """The row set of the reference's binary-fallback statement (image_database.py:1597-1605)."""
    image_ids: np.ndarray     # int64 [n], in the order the UNFILTERED statement returns rows
    codes: np.ndarray         # uint8 [n, dim] of 0/1
    file_paths: List[str]     # [n]


# The fallback's statement without its {where_clause} (image_database.py:1597-1605).
BINARY_SQL = """
    SELECT
        be.image_id,
        be.embedding,
        i.file_path
    FROM binary_embeddings be
    JOIN images i ON be.image_id = i.id
"""


def read_codes(db_path: str, expect_dim: Optional[int] = None) -> HostCodes:
    """Run the fallback's own statement (no WHERE) and keep the rows in the order SQLite
    returns them: that order is the tie-break of the reference's stable sort (:1627)."""
    conn = connect(db_path)
    try:
        cur = conn.execute(BINARY_SQL)
        ids, parts, paths = [], [], []
        dim = expect_dim
        while True:
            batch = cur.fetchmany(CHUNK_ROWS)
            if not batch:
                break
            b_ids, blobs, b_paths = zip(*batch)
            for image_id, blob in zip(b_ids, blobs):
                if dim is None:
                    dim = len(blob)
                if len(blob) != dim:
                    raise ValueError(f"binary_embeddings image_id {image_id}: {len(blob)} bytes, expected {dim}")
            ids.extend(b_ids)
            paths.extend(b_paths)
            parts.append(np.frombuffer(b"".join(blobs), dtype=np.uint8).reshape(len(blobs), dim))
    finally:
        conn.close()
    codes = np.concatenate(parts) if parts else np.zeros((0, dim or 0), dtype=np.uint8)
    return HostCodes(np.asarray(ids, dtype=np.int64), np.ascontiguousarray(codes), paths)


def connect(db_path: str, readonly: bool = True) -> sqlite3.Connection:
    if readonly:
        return sqlite3.connect(f"file:{db_path}?mode=ro", uri=True, timeout=30.0)
    return sqlite3.connect(db_path, timeout=30.0)


@contextmanager
def snapshot(db_path: str):
    """A read-only connection inside one read transaction: every statement sees the same state."""
    conn = sqlite3.connect(f"file:{db_path}?mode=ro", uri=True, timeout=30.0, isolation_level=None)
    try:
        conn.execute("BEGIN")
        yield conn
    finally:
        try:
            conn.execute("ROLLBACK")
        except sqlite3.Error:
            pass
        conn.close()


def data_version(conn: sqlite3.Connection) -> int:
    """SQLite's change counter as THIS connection sees it: differs from its previous value iff another
    connection committed in between.  O(1): what makes refresh-before-every-search affordable."""
    return int(conn.execute("PRAGMA data_version").fetchone()[0])


def _table_kind(conn: sqlite3.Connection, name: str) -> Optional[str]:
    row = conn.execute("SELECT type, sql FROM sqlite_master WHERE name = ?", (name,)).fetchone()
    if row is None:
        return None
    sql = (row[1] or "").upper()
    return "virtual" if "VIRTUAL" in sql else row[0]


def _try_load_sqlite_vec(conn: sqlite3.Connection) -> bool:
    try:
        import sqlite_vec  # noqa: F401  (absent in this image)
        conn.enable_load_extension(True)
        sqlite_vec.load(conn)
        return True
    except Exception:
        return False


def _vec0_source(conn: sqlite3.Connection) -> str:
    kind = _table_kind(conn, "vec0")
    if kind == "virtual":
        if _try_load_sqlite_vec(conn):
            return "sqlite-vec"
        if _table_kind(conn, "vec0_rowids") is None:
            raise RuntimeError("vec0 is a sqlite-vec virtual table, the extension is not importable "
                               "and its shadow tables are missing")
        return "shadow-tables"
    if kind == "table":
        return "plain-table"
    if _table_kind(conn, "vec0_rowids") is not None:
        return "shadow-tables"
    raise RuntimeError("database has no vec0 table")


def vec0_source(conn: sqlite3.Connection) -> str:
    """"plain-table", "sqlite-vec" or "shadow-tables" (raises when the database has no vec0 at all)."""
    return _vec0_source(conn)


def peek_dim(conn: sqlite3.Connection, min_rowid: Optional[int] = None, max_rowid: Optional[int] = None
             ) -> Optional[int]:
    """Dimension of the first vec0 blob with rowid in (min_rowid, max_rowid] of a plain-table vec0, or None when
    the range is empty."""
    lo = _MIN_ROWID if min_rowid is None else min_rowid
    hi = _MAX_ROWID if max_rowid is None else max_rowid
    row = conn.execute("SELECT rowid, length(embedding) FROM vec0 WHERE rowid > ? AND rowid <= ? ORDER BY rowid LIMIT 1",
                       (lo, hi)).fetchone()
    if row is None:
        return None
    if row[1] is None or row[1] % 4:
        raise ValueError(f"vec0 rowid {row[0]}: blob length {row[1]} is not float32[]")
    return int(row[1]) // 4


def count_vec0(conn: sqlite3.Connection) -> int:
    """``SELECT COUNT(*) FROM vec0`` (the reference's guard, image_database.py:1532-1540)."""
    source = _vec0_source(conn)
    table = "vec0_rowids" if source == "shadow-tables" else "vec0"
    return int(conn.execute(f"SELECT COUNT(*) FROM {table}").fetchone()[0])


def count_binary(conn: sqlite3.Connection) -> int:
    try:
        return int(conn.execute("SELECT COUNT(*) FROM binary_embeddings").fetchone()[0])
    except sqlite3.OperationalError:
        return -1   # table not accessible: the reference returns [] (:1496-1500)


def _iter_vec0_batches(conn: sqlite3.Connection, source: str, lo: int, hi: int, chunk_rows: int
                       ) -> Iterator[Tuple[np.ndarray, Sequence[bytes]]]:
    """(rowids int64 [m], m float32 blobs) in ascending rowid order for rowids in (lo, hi]."""
    if source in ("sqlite-vec", "plain-table"):
        cur = conn.execute("SELECT rowid, embedding FROM vec0 WHERE rowid > ? AND rowid <= ? ORDER BY rowid", (lo, hi))
        while True:
            batch = cur.fetchmany(chunk_rows)
            if not batch:
                return
            ids, blobs = zip(*batch)
            yield np.asarray(ids, dtype=np.int64), blobs
    else:
        yield from _iter_shadow_batches(conn, lo, hi, chunk_rows)


def _iter_shadow_batches(conn: sqlite3.Connection, lo: int, hi: int, chunk_rows: int
                         ) -> Iterator[Tuple[np.ndarray, Sequence[bytes]]]:
    """Walk sqlite-vec's chunked storage in rowid order: one ``vec0_rowids`` range at a time, each storage
    chunk fetched once and sliced with numpy."""
    cache = {}

    def chunk(cid: int):
        if cid not in cache:
            if len(cache) > 16:
                cache.clear()
            size, validity = conn.execute("SELECT size, validity FROM vec0_chunks WHERE chunk_id = ?", (cid,)).fetchone()
            vectors = conn.execute("SELECT vectors FROM vec0_vector_chunks00 WHERE rowid = ?", (cid,)).fetchone()[0]
            valid = np.unpackbits(np.frombuffer(validity, dtype=np.uint8), bitorder="little").astype(bool)
            cache[cid] = (int(size), valid, memoryview(vectors), len(vectors) // int(size))
        return cache[cid]

    cur = conn.execute("SELECT rowid, chunk_id, chunk_offset FROM vec0_rowids WHERE rowid > ? AND rowid <= ? "
                       "ORDER BY rowid", (lo, hi))
    while True:
        batch = cur.fetchmany(chunk_rows)
        if not batch:
            return
        ids, blobs = [], []
        for rowid, cid, off in batch:
            size, valid, vectors, stride = chunk(cid)
            if off >= size or not valid[off]:
                continue
            ids.append(rowid)
            blobs.append(vectors[off * stride:(off + 1) * stride])
        if ids:
            yield np.asarray(ids, dtype=np.int64), blobs


def _mapping_range(conn: sqlite3.Connection, first: int, last: int, with_paths: bool = True):
    """``image_embeddings JOIN images`` for rowids in [first, last], ascending: (rowids, image_ids, mtimes, paths)."""
    cols = "ie.rowid, ie.image_id, i.last_modified" + (", i.file_path" if with_paths else "")
    got = conn.execute(f"SELECT {cols} FROM image_embeddings ie JOIN images i ON ie.image_id = i.id "
                       "WHERE ie.rowid >= ? AND ie.rowid <= ? ORDER BY ie.rowid", (first, last)).fetchall()
    if not got:
        return (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float64), [])
    cols_t = list(zip(*got))
    mt = np.asarray([0.0 if v is None else v for v in cols_t[2]], dtype=np.float64)
    return (np.asarray(cols_t[0], dtype=np.int64), np.asarray(cols_t[1], dtype=np.int64), mt,
            list(cols_t[3]) if with_paths else [])


def iter_store(conn: sqlite3.Connection, min_rowid: Optional[int] = None, max_rowid: Optional[int] = None,
               chunk_rows: int = CHUNK_ROWS, expect_dim: Optional[int] = None, stats: Optional[StoreStats] = None
               ) -> Iterator[StoreChunk]:
    """The rows the search statement scans, for rowids in (min_rowid, max_rowid], chunk by chunk in scan order."""
    st = stats if stats is not None else StoreStats()
    st.source = _vec0_source(conn)
    st.dim = expect_dim
    lo = _MIN_ROWID if min_rowid is None else min_rowid
    hi = _MAX_ROWID if max_rowid is None else max_rowid
    for ids, blobs in _iter_vec0_batches(conn, st.source, lo, hi, chunk_rows):
        st.vec0_rows += len(ids)
        m_ids, m_image, m_mtime, m_paths = _mapping_range(conn, int(ids[0]), int(ids[-1]))
        if len(m_ids) == 0:
            continue
        at = np.searchsorted(m_ids, ids)
        at[at >= len(m_ids)] = len(m_ids) - 1
        keep = m_ids[at] == ids
        if not keep.all():
            if not keep.any():
                continue
            blobs = [b for b, k in zip(blobs, keep.tolist()) if k]
            ids, at = ids[keep], at[keep]
        lengths = set(map(len, blobs))
        if st.dim is None:
            first_len = len(blobs[0])
            if first_len % 4:
                raise ValueError(f"vec0 rowid {int(ids[0])}: blob length {first_len} is not float32[]")
            st.dim = first_len // 4
        if lengths != {st.dim * 4}:
            for rowid, blob in zip(ids.tolist(), blobs):
                if len(blob) % 4:
                    raise ValueError(f"vec0 rowid {rowid}: blob length {len(blob)} is not float32[]")
                if len(blob) != st.dim * 4:
                    raise ValueError(f"vec0 rowid {rowid}: {len(blob) // 4} floats, expected {st.dim}")
        st.joined_rows += len(ids)
        whole = len(at) == len(m_ids)
        yield StoreChunk(ids, m_image if whole else m_image[at], m_mtime if whole else m_mtime[at],
                         m_paths if whole else [m_paths[i] for i in at.tolist()], blobs=blobs, dim=st.dim)


def read_mapping(conn: sqlite3.Connection, min_rowid: Optional[int] = None, max_rowid: Optional[int] = None
                 ) -> Mapping:
    """Every ``image_embeddings`` row that has its ``images`` partner, for rowids in (min_rowid, max_rowid]:
    integers and one float per row, no blobs, no strings."""
    lo = _MIN_ROWID if min_rowid is None else min_rowid
    hi = _MAX_ROWID if max_rowid is None else max_rowid
    cur = conn.execute("SELECT ie.rowid, ie.image_id, i.last_modified FROM image_embeddings ie "
                       "JOIN images i ON ie.image_id = i.id WHERE ie.rowid > ? AND ie.rowid <= ? ORDER BY ie.rowid",
                       (lo, hi))
    ids, images, mtimes = [], [], []
    while True:
        batch = cur.fetchmany(1 << 16)
        if not batch:
            break
        a, b, c = zip(*batch)
        ids.append(np.asarray(a, dtype=np.int64))
        images.append(np.asarray(b, dtype=np.int64))
        mtimes.append(np.asarray([0.0 if v is None else v for v in c], dtype=np.float64))
    if not ids:
        z = np.zeros(0, dtype=np.int64)
        return Mapping(z, z.copy(), np.zeros(0, dtype=np.float64))
    return Mapping(np.concatenate(ids), np.concatenate(images), np.concatenate(mtimes))


def plan_shards(conn: sqlite3.Connection, world: int) -> Tuple[List[Tuple[Optional[int], Optional[int]]], int]:
    """Row-sharding of a real database: the mapped rowids (``image_embeddings JOIN images``), ascending, cut
    into ``world`` contiguous ranges of equal size (+-1).  Returns ([(min_rowid exclusive, max_rowid
    inclusive)] per rank, mapped rows): the ranges tile the whole rowid axis, so every vec0 row belongs to
    exactly one of them, and they come from ONE statement on the caller's connection (hold it in a
    ``snapshot`` and read every range through the same connection)."""
    ids = read_mapping(conn).rowids
    n = len(ids)
    ranges: List[Tuple[Optional[int], Optional[int]]] = []
    for rank in range(world):
        lo, hi = n * rank // world, n * (rank + 1) // world
        if hi <= lo:
            ranges.append((None, None))
            continue
        ranges.append((int(ids[lo - 1]) if lo > 0 else None, int(ids[hi - 1]) if rank < world - 1 else None))
    return ranges, n


def shard_rowid_range(db_path: str, rank: int, world: int) -> Tuple[Optional[int], Optional[int], int]:
    """One rank's (min_rowid exclusive, max_rowid inclusive, mapped rows); see ``plan_shards``.  The last
    rank's upper bound is reported as its last mapped rowid."""
    with snapshot(db_path) as conn:
        ids = read_mapping(conn).rowids
    n = len(ids)
    lo, hi = n * rank // world, n * (rank + 1) // world
    if hi <= lo:
        return None, None, n
    return (int(ids[lo - 1]) if lo > 0 else None), int(ids[hi - 1]), n


def read_rows_by_rowid(conn: sqlite3.Connection, rowids: Sequence[int]) -> Iterator[StoreChunk]:
    """The current blobs and mapping of specific rowids (refresh after an in-place re-embedding,
    image_database.py:1165-1167): one range read per run of nearby rowids."""
    want = np.unique(np.asarray(list(rowids), dtype=np.int64))
    i = 0
    while i < len(want):
        j = i
        while j + 1 < len(want) and want[j + 1] - want[i] < CHUNK_ROWS:
            j += 1
        for chunk in iter_store(conn, int(want[i]) - 1, int(want[j])):
            keep = np.isin(chunk.rowids, want[i:j + 1])
            if keep.any():
                idx = np.flatnonzero(keep)
                yield StoreChunk(chunk.rowids[idx], chunk.image_ids[idx], chunk.mtimes[idx],
                                 [chunk.file_paths[t] for t in idx.tolist()], rows=chunk.rows[idx])
        i = j + 1


def stream_store(db_path: str, sink, expect_dim: Optional[int] = None, min_rowid: Optional[int] = None,
                 max_rowid: Optional[int] = None, chunk_rows: int = CHUNK_ROWS,
                 conn: Optional[sqlite3.Connection] = None) -> HostStore:
    """Feed every chunk to ``sink(chunk)`` (which copies what it needs, e.g. ``chunk.write_rows(buffer)``) and
    return the store's metadata with ``rows=None``."""
    if conn is None:
        with snapshot(db_path) as own:
            return stream_store(db_path, sink, expect_dim, min_rowid, max_rowid, chunk_rows, own)
    st = StoreStats()
    ids, images, mtimes, paths = [], [], [], []
    for chunk in iter_store(conn, min_rowid, max_rowid, chunk_rows, expect_dim, st):
        sink(chunk)
        ids.append(chunk.rowids)
        images.append(chunk.image_ids)
        mtimes.append(chunk.mtimes)
        paths.extend(chunk.file_paths)
    cat = lambda parts, dt: np.concatenate(parts) if parts else np.zeros(0, dtype=dt)   # noqa: E731
    return HostStore(cat(ids, np.int64), None, cat(images, np.int64), paths, count_binary(conn), st.vec0_rows,
                     st.source, st.vec0_rows - st.joined_rows, cat(mtimes, np.float64), st.dim or 0)


def read_store(db_path: str, expect_dim: Optional[int] = None, min_rowid: Optional[int] = None,
               max_rowid: Optional[int] = None) -> HostStore:
    """Everything the resident index needs, in scan order, as host arrays (tests, small databases; the
    product path streams, see ``stream_store``).  ``min_rowid`` restricts the read to rowids greater than it
    (incremental refresh after the scanner appended rows; the reference never deletes from vec0);
    ``max_rowid`` (inclusive) bounds it from above (one rank's range of a row-sharded store)."""
    with snapshot(db_path) as conn:
        lo = _MIN_ROWID if min_rowid is None else min_rowid
        hi = _MAX_ROWID if max_rowid is None else max_rowid
        bound = int(conn.execute("SELECT COUNT(*) FROM image_embeddings WHERE rowid > ? AND rowid <= ?",
                                 (lo, hi)).fetchone()[0])
        box = {"rows": None, "n": 0}

        def sink(chunk: StoreChunk) -> None:
            if box["rows"] is None:
                box["rows"] = np.empty((bound, chunk.dim), dtype=np.float32)
            m = len(chunk)
            chunk.write_rows(box["rows"][box["n"]:box["n"] + m])
            box["n"] += m
        host = stream_store(db_path, sink, expect_dim, min_rowid, max_rowid, CHUNK_ROWS, conn)
    rows = box["rows"]
    if rows is None:
        rows = np.zeros((0, host.dim_), dtype=np.float32)
    elif box["n"] != rows.shape[0]:
        rows = np.ascontiguousarray(rows[:box["n"]])
    host.rows = rows
    return host
