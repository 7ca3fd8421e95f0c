"""Seeded synthetic data in the reference's formats.

``unit_rows`` is the numpy generator used for host-side fixtures (config 1);
``write_reference_db`` lays rows out in the reference's SQLite schema the way
``_commit_batch`` does (image_database.py:1137-1198): one ``images`` row, one
``vec0`` row whose rowid is mirrored into ``image_embeddings``, and one 1152-byte
sign code in ``binary_embeddings`` per image.
"""
from __future__ import annotations

import os
import sqlite3
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import schema


def unit_rows(n: int, dim: int = schema.EMBEDDING_DIM, seed: int = 1234) -> np.ndarray:
    """``default_rng(seed).standard_normal`` float32 rows, L2-normalised in float32."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, dim), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
    return x


def fp16_normalised(v: np.ndarray) -> np.ndarray:
    """What the reference stores on CUDA: normalised in float16, then widened to
    float32 (image_database.py:457-458, 493-494), so ||v|| is only ~0.9999."""
    h = np.asarray(v, dtype=np.float16)
    h = h / np.linalg.norm(h.astype(np.float32), axis=-1, keepdims=True).astype(np.float16)
    return h.astype(np.float32)


def default_paths(n: int, folders: Sequence[str] = ("/data/photos/a", "/data/photos/b", "/data/scans")
                  ) -> List[str]:
    return [f"{folders[i % len(folders)]}/img_{i:08d}.jpg" for i in range(n)]


def write_reference_db(path: str, rows: np.ndarray, file_paths: Optional[Sequence[str]] = None,
                       vec0_layout: str = "standin", rowid_start: int = 1,
                       drop_mapping_for: Iterable[int] = (), drop_image_for: Iterable[int] = (),
                       binary_codes: bool = True, chunk_size: int = 1024, vectors: bool = True) -> None:
    """Create ``path`` with ``rows[i]`` stored as vec0 rowid ``rowid_start + i``.

    ``drop_mapping_for`` / ``drop_image_for`` are positions whose ``image_embeddings``
    / ``images`` row is removed afterwards: orphaned vec0 rows that the reference's
    INNER JOINs silently exclude (image_database.py:1569-1570).
    ``vec0_layout``: "standin" (plain table) or "shadow" (sqlite-vec shadow tables).
    ``vectors=False`` leaves ``vec0`` and ``image_embeddings`` empty: a binary-only database,
    which makes the reference's search() take its sign-code fallback (image_database.py:1591-1629).
    """
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n, dim = rows.shape
    if file_paths is None:
        file_paths = default_paths(n)
    if os.path.exists(path):
        os.remove(path)
    conn = sqlite3.connect(path)
    cur = conn.cursor()
    cur.execute(schema.IMAGES)
    cur.execute(schema.IMAGE_EMBEDDINGS)
    cur.execute(schema.BINARY_EMBEDDINGS)
    cur.execute(schema.BINARY_EMBEDDINGS_INDEX)

    cur.executemany("INSERT INTO images (id, file_path, last_modified, file_hash) VALUES (?, ?, ?, ?)",
                    ((i + 1, file_paths[i], 1.7e9 + i, f"{i:032x}") for i in range(n)))
    rowids = np.arange(rowid_start, rowid_start + n, dtype=np.int64)
    if not vectors:
        cur.execute(schema.VEC0_STANDIN)
    elif vec0_layout == "standin":
        cur.execute(schema.VEC0_STANDIN)
        cur.executemany("INSERT INTO vec0 (rowid, embedding) VALUES (?, ?)",
                        ((int(rowids[i]), rows[i].tobytes()) for i in range(n)))
    elif vec0_layout == "shadow":
        cur.execute(schema.SHADOW_CHUNKS)
        cur.execute(schema.SHADOW_ROWIDS)
        cur.execute(schema.SHADOW_VECTORS)
        for c, lo in enumerate(range(0, n, chunk_size)):
            hi = min(lo + chunk_size, n)
            m = hi - lo
            valid = np.zeros(chunk_size, dtype=bool)
            valid[:m] = True
            ids = np.zeros(chunk_size, dtype=np.int64)
            ids[:m] = rowids[lo:hi]
            vec = np.zeros((chunk_size, dim), dtype=np.float32)
            vec[:m] = rows[lo:hi]
            cur.execute("INSERT INTO vec0_chunks (chunk_id, size, validity, rowids) VALUES (?, ?, ?, ?)",
                        (c + 1, chunk_size, np.packbits(valid, bitorder="little").tobytes(), ids.tobytes()))
            cur.execute("INSERT INTO vec0_vector_chunks00 (rowid, vectors) VALUES (?, ?)",
                        (c + 1, vec.tobytes()))
            cur.executemany("INSERT INTO vec0_rowids (rowid, id, chunk_id, chunk_offset) VALUES (?, NULL, ?, ?)",
                            ((int(rowids[lo + j]), c + 1, j) for j in range(m)))
    else:
        raise ValueError("vec0_layout must be 'standin' or 'shadow'")
    if vectors:
        cur.executemany("INSERT INTO image_embeddings (rowid, image_id) VALUES (?, ?)",
                        ((int(rowids[i]), i + 1) for i in range(n)))
    if binary_codes:
        codes = (rows >= 0).astype(np.uint8)  # image_database.py:1189-1190
        cur.executemany("INSERT INTO binary_embeddings (image_id, embedding) VALUES (?, ?)",
                        ((i + 1, codes[i].tobytes()) for i in range(n)))
    for pos in drop_mapping_for:
        cur.execute("DELETE FROM image_embeddings WHERE rowid = ?", (int(rowids[pos]),))
    for pos in drop_image_for:
        cur.execute("DELETE FROM images WHERE id = ?", (pos + 1,))
    conn.commit()
    conn.close()
