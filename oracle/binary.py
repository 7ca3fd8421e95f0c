"""``oracle_binary`` — CPU restatement of the reference's sign-code fallback search.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): only tests/, ``smoke()`` and bench.py's
CPU leg may import this.

Follows image_database.py:1591-1629: ``query_binary = (embedding >= 0).astype(np.uint8)``
(:1593); per fetched row ``binary_score = np.dot(query_binary, candidate_binary)`` (:1621) —
evaluated by numpy IN uint8, i.e. modulo 256 (pinned by ``tests/golden`` cases produced by the
reference's own code); ``similarity = float(binary_score) / 1152`` (:1624); a stable
descending sort on similarity (:1627) and ``candidate_scores[:k]`` (:1628) — a Python slice,
so a negative k drops the last |k| rows instead of meaning "no limit".
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def sign_code(embedding: np.ndarray) -> np.ndarray:
    """image_database.py:1593 / :1189."""
    return (np.asarray(embedding) >= 0).astype(np.uint8)


def scores(codes: np.ndarray, query_code: np.ndarray, wrap: bool = True) -> np.ndarray:
    """AND-popcount per row; ``wrap`` reproduces the uint8 arithmetic of np.dot(uint8, uint8)."""
    c = np.asarray(codes, dtype=np.uint8)
    q = np.asarray(query_code, dtype=np.uint8).astype(np.int32)
    s = np.empty(c.shape[0], dtype=np.int64)
    for lo in range(0, c.shape[0], 16384):          # exact integer dot product, bounded temporaries
        s[lo:lo + 16384] = c[lo:lo + 16384].astype(np.int32) @ q
    return (s % 256) if wrap else s


def search(codes: np.ndarray, query_code: np.ndarray, k: int, wrap: bool = True,
           order: Optional[Sequence[int]] = None) -> Tuple[np.ndarray, np.ndarray]:
    """(positions, scores) of the reference's top-k.  ``order``: positions in the order the
    statement returned them (all rows in scan order when None); rows not listed were filtered."""
    s = scores(codes, query_code, wrap)
    seq = np.arange(len(s)) if order is None else np.asarray(order, dtype=np.int64)
    ranked = sorted(((int(p), float(s[p]) / 1152.0) for p in seq), key=lambda x: x[1], reverse=True)
    top = ranked[:k]
    pos = np.array([p for p, _ in top], dtype=np.int64)
    return pos, s[pos] if len(pos) else np.zeros(0, dtype=np.int64)
