"""``oracle_sql`` — the reference's search statement run by the real SQLite.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED for the
distance arithmetic (third-party sqlite-vec, absent); everything else here —
the joins, the WHERE pre-filter, ORDER BY/LIMIT admission and tie order, NULL
ordering — is executed by SQLite 3.45's own code on the reference's SQL text,
so those semantics are pinned by the real thing rather than restated.

``vec_distance_cosine`` is provided by ``vec_shim.so`` (a loadable extension
wrapping oracle_ref.c), or by ``sqlite_vec`` itself should it ever import.
"""
from __future__ import annotations

import os
import sqlite3
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import ref

# The statement at image_database.py:1564-1574, with its {where_clause} slot.
SEARCH_SQL = """
    SELECT
        i.file_path,
        vec_distance_cosine(vec0.embedding, ?) as distance
    FROM vec0
    JOIN image_embeddings ie ON vec0.rowid = ie.rowid
    JOIN images i ON ie.image_id = i.id
    {where_clause}
    ORDER BY distance ASC
    LIMIT ?
"""

# Same statement with the rowid exposed, for id-level parity checks.
SEARCH_SQL_WITH_ROWID = SEARCH_SQL.replace("i.file_path,", "i.file_path, vec0.rowid,")


def connect(db_path: str) -> Tuple[sqlite3.Connection, str]:
    """Connection with ``vec_distance_cosine`` registered, as after
    ``sqlite_vec.load(conn)`` (image_database.py:1475-1484).  Returns (conn, provider)."""
    conn = sqlite3.connect(db_path, timeout=30.0)
    conn.execute("PRAGMA journal_mode=WAL")
    conn.enable_load_extension(True)
    try:
        import sqlite_vec  # absent in this image
        sqlite_vec.load(conn)
        return conn, "sqlite-vec"
    except Exception:
        pass
    ref.build()
    try:
        conn.load_extension(ref.SHIM_PATH)
        return conn, "vec_shim (oracle_ref.c)"
    except sqlite3.OperationalError:
        def dist(a: bytes, b: bytes):
            d = ref.cosine_distance(np.frombuffer(a, dtype=np.float32), np.frombuffer(b, dtype=np.float32))
            return d  # a NaN float becomes SQL NULL, as with sqlite3_result_double
        conn.create_function("vec_distance_cosine", 2, dist, deterministic=True)
        return conn, "python callback (oracle_ref.c via ctypes)"


def where_clause_and_params(filter_folders: Optional[Sequence[str]]) -> Tuple[str, List[str]]:
    """image_database.py:1509-1530 and :1576-1579."""
    if not filter_folders:
        return "", []
    normalized = []
    for folder in filter_folders:
        folder_abs = os.path.abspath(folder)
        if not folder_abs.endswith(os.sep):
            folder_abs += os.sep
        normalized.append(folder_abs)
    conds = ["i.file_path LIKE ? ESCAPE '\\'" for _ in normalized]
    params = [f.replace("\\", "\\\\").replace("%", "\\%").replace("_", "\\_") + "%" for f in normalized]
    return "WHERE (" + " OR ".join(conds) + ")", params


def run_statement(conn: sqlite3.Connection, embedding: np.ndarray, k: int,
                  filter_folders: Optional[Sequence[str]] = None, with_rowid: bool = False):
    """Execute the search statement; raw rows ((file_path[, rowid], distance))."""
    where, fparams = where_clause_and_params(filter_folders)
    sql = (SEARCH_SQL_WITH_ROWID if with_rowid else SEARCH_SQL).format(where_clause=where)
    query_vec = np.asarray(embedding, dtype=np.float32).tobytes()  # == serialize_float32 (:1561)
    cur = conn.cursor()
    cur.execute(sql, tuple([query_vec] + fparams + [k]))
    return cur.fetchall()


def reference_search(db_path: str, embedding: np.ndarray, k: int,
                     filter_folders: Optional[Sequence[str]] = None) -> List[Tuple[str, float]]:
    """The db part of ``search()`` with its guards and error envelope
    (image_database.py:1474-1642), before the duplicate filter."""
    conn, _ = connect(db_path)
    cur = conn.cursor()
    try:
        cur.execute("SELECT COUNT(*) FROM binary_embeddings")
        if cur.fetchone()[0] == 0:
            conn.close()
            return []
    except sqlite3.OperationalError:
        conn.close()
        return []
    try:
        cur.execute("SELECT COUNT(*) FROM vec0")
        if cur.fetchone()[0] <= 0:
            raise NotImplementedError("binary fallback path is out of scope")
        rows = run_statement(conn, embedding, k, filter_folders)
        top = []
        for file_path, distance in rows:
            top.append((file_path, 1.0 - distance))   # raises on NULL, like the reference
        results = [(p, float(s)) for p, s in top]
    except NotImplementedError:
        conn.close()
        raise
    except Exception:
        conn.close()
        return []
    conn.close()
    return results


# The fallback's statement (image_database.py:1597-1605) with its {where_clause} slot.
BINARY_SQL = """
    SELECT
        be.image_id,
        be.embedding,
        i.file_path
    FROM binary_embeddings be
    JOIN images i ON be.image_id = i.id
    {where_clause}
"""


def reference_binary_search(db_path: str, embedding: np.ndarray, k: int,
                            filter_folders: Optional[Sequence[str]] = None) -> List[Tuple[str, float]]:
    """The sign-code fallback of ``search()`` written the literal way (image_database.py:1591-1629):
    the statement executed by the real SQLite (so the arrival order, which breaks ties in the
    stable sort, is SQLite's own), then the reference's per-row numpy arithmetic."""
    conn = sqlite3.connect(db_path, timeout=30.0)
    where, fparams = where_clause_and_params(filter_folders)
    query_binary = (np.asarray(embedding) >= 0).astype(np.uint8)
    rows = conn.execute(BINARY_SQL.format(where_clause=where), tuple(fparams)).fetchall()
    conn.close()
    candidate_scores = []
    for _image_id, binary_blob, file_path in rows:
        candidate_binary = np.frombuffer(binary_blob, dtype=np.uint8)
        binary_score = np.dot(query_binary, candidate_binary)      # uint8 arithmetic: modulo 256
        candidate_scores.append((file_path, float(binary_score) / query_binary.shape[0]))
    candidate_scores.sort(key=lambda x: x[1], reverse=True)
    return candidate_scores[:k]


def reference_filter_duplicates(db_path: str, results: List[Tuple[str, float]],
                                tolerance_bits: int = 2) -> List[Tuple[str, float]]:
    """Behavioural restatement of ``_filter_duplicates`` (image_database.py:1207-1306),
    deliberately written the slow, literal way (dict of tuples, per-pair compare)."""
    if len(results) == 0:
        return results
    conn = sqlite3.connect(db_path, timeout=30.0)
    cur = conn.cursor()
    file_to_id, id_to_binary = {}, {}
    for file_path, _ in results:
        row = cur.execute("SELECT id FROM images WHERE file_path = ?", (file_path,)).fetchone()
        if row:
            file_to_id[file_path] = row[0]
    if file_to_id:
        ids = list(file_to_id.values())
        marks = ",".join(["?"] * len(ids))
        for image_id, blob in cur.execute(
                f"SELECT image_id, embedding FROM binary_embeddings WHERE image_id IN ({marks})", ids):
            id_to_binary[image_id] = np.frombuffer(blob, dtype=np.uint8)
    conn.close()
    seen = {}
    filtered = []
    for file_path, similarity in results:
        image_id = file_to_id.get(file_path)
        if image_id is None or image_id not in id_to_binary:
            filtered.append((file_path, similarity))
            continue
        code = id_to_binary[image_id]
        duplicate = False
        for seen_key, (seen_path, seen_sim) in seen.items():
            if int(np.sum(code != np.array(seen_key, dtype=np.uint8))) <= tolerance_bits:
                duplicate = True
                if similarity > seen_sim:
                    seen[seen_key] = (file_path, similarity)
                    filtered = [(fp, s) for fp, s in filtered if fp != seen_path]
                    filtered.append((file_path, similarity))
                break
        if not duplicate:
            seen[tuple(code)] = (file_path, similarity)
            filtered.append((file_path, similarity))
    filtered.sort(key=lambda x: x[1], reverse=True)
    return filtered
