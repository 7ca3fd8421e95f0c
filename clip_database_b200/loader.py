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
"""
from __future__ import annotations

import sqlite3
from dataclasses import dataclass
from typing import Iterator, List, Optional, Tuple

import numpy as np


@dataclass
class HostStore:
    rowids: np.ndarray        # int64 [n], ascending (= scan order)
    rows: np.ndarray          # float32 [n, dim]
    image_ids: np.ndarray     # int64 [n]
    file_paths: List[str]     # [n]
    binary_count: int         # COUNT(*) FROM binary_embeddings (guard, image_database.py:1488-1500)
    vec0_count: int           # COUNT(*) FROM vec0 before the joins (guard, :1532-1540)
    source: str               # which of the three vec0 readers was used
    dropped: int              # vec0 rows excluded by the INNER JOINs

    @property
    def dim(self) -> int:
        return int(self.rows.shape[1]) if self.rows.ndim == 2 else 0


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
        rows = conn.execute(BINARY_SQL).fetchall()
    finally:
        conn.close()
    dim = expect_dim
    for image_id, blob, _ in rows:
        if dim is None:
            dim = len(blob)
        if len(blob) != dim:
            raise ValueError(f"binary_embeddings image_id {image_id}: {len(blob)} bytes, expected {dim}")
    n = len(rows)
    codes = np.frombuffer(b"".join(r[1] for r in rows), dtype=np.uint8).reshape(n, dim or 0)
    return HostCodes(np.asarray([r[0] for r in rows], dtype=np.int64), np.ascontiguousarray(codes),
                     [r[2] for r in rows])


def connect(db_path: str, readonly: bool = True) -> sqlite3.Connection:
    if readonly:
        return sqlite3.connect(f"file:{db_path}?mode=ro", uri=True, timeout=30.0)
    return sqlite3.connect(db_path, timeout=30.0)


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


def _iter_vec0(conn: sqlite3.Connection, min_rowid: Optional[int], max_rowid: Optional[int] = None
               ) -> Tuple[str, Iterator[Tuple[int, bytes]]]:
    """(source, iterator of (rowid, float32 blob)) in ascending rowid order, for rowids in
    (min_rowid, max_rowid]."""
    kind = _table_kind(conn, "vec0")
    lo = -(1 << 63) if min_rowid is None else min_rowid
    hi = (1 << 63) - 1 if max_rowid is None else max_rowid
    if kind == "virtual":
        if _try_load_sqlite_vec(conn):
            cur = conn.execute("SELECT rowid, embedding FROM vec0 WHERE rowid > ? AND rowid <= ? ORDER BY rowid", (lo, hi))
            return "sqlite-vec", iter(cur)
        if _table_kind(conn, "vec0_rowids") is None:
            raise RuntimeError("vec0 is a sqlite-vec virtual table, the extension is not importable "
                               "and its shadow tables are missing")
        return "shadow-tables", _iter_shadow(conn, lo, hi)
    if kind == "table":
        cur = conn.execute("SELECT rowid, embedding FROM vec0 WHERE rowid > ? AND rowid <= ? ORDER BY rowid", (lo, hi))
        return "plain-table", iter(cur)
    if _table_kind(conn, "vec0_rowids") is not None:
        return "shadow-tables", _iter_shadow(conn, lo, hi)
    raise RuntimeError("database has no vec0 table")


def _vec0_rowids(conn: sqlite3.Connection) -> List[int]:
    """Every vec0 rowid (no blobs), ascending."""
    kind = _table_kind(conn, "vec0")
    if kind == "table" or (kind == "virtual" and _try_load_sqlite_vec(conn)):
        return [r[0] for r in conn.execute("SELECT rowid FROM vec0 ORDER BY rowid")]
    if _table_kind(conn, "vec0_rowids") is not None:
        return [r[0] for r in conn.execute("SELECT rowid FROM vec0_rowids ORDER BY rowid")]
    raise RuntimeError("database has no vec0 table")


def shard_rowid_range(db_path: str, rank: int, world: int) -> Tuple[Optional[int], Optional[int], int]:
    """Row-sharding of a real database: the rows the search statement scans (vec0 INNER JOIN
    image_embeddings INNER JOIN images), in rowid order, cut into `world` contiguous ranges of
    equal size (+-1).  Returns (min_rowid exclusive, max_rowid inclusive, total joined rows) for
    `rank` — only that range's blobs need to be read by that rank."""
    conn = connect(db_path)
    try:
        partner = {r[0] for r in conn.execute(
            "SELECT ie.rowid FROM image_embeddings ie JOIN images i ON ie.image_id = i.id")}
        joined = [r for r in _vec0_rowids(conn) if r in partner]
    finally:
        conn.close()
    n = len(joined)
    lo, hi = n * rank // world, n * (rank + 1) // world
    if hi <= lo:
        return None, None, n
    return (joined[lo - 1] if lo > 0 else None), joined[hi - 1], n


def _iter_shadow(conn: sqlite3.Connection, lo: int, hi: int = (1 << 63) - 1) -> Iterator[Tuple[int, bytes]]:
    """Walk sqlite-vec's chunked storage in rowid order."""
    chunk_cache = {}

    def chunk(cid: int):
        if cid not in chunk_cache:
            if len(chunk_cache) > 8:
                chunk_cache.clear()
            size, validity = conn.execute(
                "SELECT size, validity FROM vec0_chunks WHERE chunk_id = ?", (cid,)).fetchone()
            vectors = conn.execute(
                "SELECT vectors FROM vec0_vector_chunks00 WHERE rowid = ?", (cid,)).fetchone()[0]
            valid = np.unpackbits(np.frombuffer(validity, dtype=np.uint8), bitorder="little")
            chunk_cache[cid] = (size, valid, memoryview(vectors))
        return chunk_cache[cid]

    cur = conn.execute("SELECT rowid, chunk_id, chunk_offset FROM vec0_rowids WHERE rowid > ? AND rowid <= ? "
                       "ORDER BY rowid", (lo, hi))
    for rowid, cid, off in cur.fetchall():
        size, valid, vectors = chunk(cid)
        if off >= size or not valid[off]:
            continue
        stride = len(vectors) // size
        yield rowid, bytes(vectors[off * stride:(off + 1) * stride])


def read_store(db_path: str, expect_dim: Optional[int] = None, min_rowid: Optional[int] = None,
               max_rowid: Optional[int] = None) -> HostStore:
    """Everything the resident index needs, in scan order.  ``min_rowid`` restricts
    the read to rowids greater than it (incremental refresh after the scanner
    appended rows; the reference never deletes from vec0); ``max_rowid`` (inclusive) bounds it
    from above (one rank's range of a row-sharded store, see ``shard_rowid_range``)."""
    conn = connect(db_path)
    try:
        try:
            binary_count = conn.execute("SELECT COUNT(*) FROM binary_embeddings").fetchone()[0]
        except sqlite3.OperationalError:
            binary_count = -1   # table not accessible: the reference returns [] (:1496-1500)
        join = conn.execute(
            "SELECT ie.rowid, ie.image_id, i.file_path FROM image_embeddings ie "
            "JOIN images i ON ie.image_id = i.id WHERE ie.rowid > ? AND ie.rowid <= ? ORDER BY ie.rowid",
            (-(1 << 63) if min_rowid is None else min_rowid,
             (1 << 63) - 1 if max_rowid is None else max_rowid)).fetchall()
        partner = {r[0]: (r[1], r[2]) for r in join}
        source, it = _iter_vec0(conn, min_rowid, max_rowid)
        rowids, image_ids, paths, blobs = [], [], [], []
        vec0_count = 0
        dim = expect_dim
        for rowid, blob in it:
            vec0_count += 1
            hit = partner.get(rowid)
            if hit is None:
                continue
            if len(blob) % 4:
                raise ValueError(f"vec0 rowid {rowid}: blob length {len(blob)} is not float32[]")
            if dim is None:
                dim = len(blob) // 4
            if len(blob) != dim * 4:
                raise ValueError(f"vec0 rowid {rowid}: {len(blob) // 4} floats, expected {dim}")
            rowids.append(rowid)
            image_ids.append(hit[0])
            paths.append(hit[1])
            blobs.append(blob)
        n = len(rowids)
        rows = np.frombuffer(b"".join(blobs), dtype="<f4").reshape(n, dim or 0).astype(np.float32, copy=False)
        return HostStore(np.asarray(rowids, dtype=np.int64), np.ascontiguousarray(rows),
                         np.asarray(image_ids, dtype=np.int64), paths, binary_count, vec0_count,
                         source, vec0_count - n)
    finally:
        conn.close()
