"""numpy restatement of the reference's query arithmetic — TEST INFRASTRUCTURE.

Follows image_database.py (``idb``) line for line in *behaviour*, executed by
the same numpy float32 kernels the reference uses, so the GPU blend kernel is
diffed against the real thing for this part of the path (numpy is first-party
to the reference here; only ``np.linalg.norm``'s BLAS summation order is
platform-defined, hence a float tolerance rather than bit equality).

  idb:1378-1383  weight normalisation            -> ``normalise_weights``
  idb:1387-1395  positive blend + L2 normalise   -> ``blend_positive``
  idb:545-571    single negative                 -> ``apply_negative``
  idb:573-604    several negatives, list order   -> ``apply_negatives``
  idb:1402-1472  dispatch 1-vs-many              -> ``compose_query``
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def normalise_weights(weights: Tuple[float, float]) -> Tuple[Tuple[float, float], float, float]:
    """idb:1378-1383.  Returns (possibly replaced weights, w1, w2) as Python floats."""
    total = weights[0] + weights[1]
    if total == 0:
        weights = (0.5, 0.5)
        total = 1.0
    return weights, weights[0] / total, weights[1] / total


def blend_positive(e1: np.ndarray, e2: np.ndarray, weights: Tuple[float, float]
                   ) -> Tuple[np.ndarray, Tuple[float, float]]:
    """idb:1378-1395: weighted sum in float32, L2 normalise, zero norm -> e1."""
    weights, w1, w2 = normalise_weights(weights)
    mixed = w1 * e1 + w2 * e2            # python-float * float32 array stays float32
    nrm = np.linalg.norm(mixed)
    if nrm > 0:
        mixed = mixed / nrm
    else:
        mixed = e1
    return mixed, weights


def _restore(e1: np.ndarray, e2: Optional[np.ndarray], weights: Tuple[float, float]) -> np.ndarray:
    """Zero-norm fallback shared by idb:563-570 and idb:596-603."""
    if e2 is None:
        return e1
    s = weights[0] + weights[1]
    w1, w2 = weights[0] / s, weights[1] / s
    out = w1 * e1 + w2 * e2
    nrm = np.linalg.norm(out)
    if nrm > 0:
        out = out / nrm
    return out


def apply_negative(e: np.ndarray, neg: np.ndarray, neg_w: float, e1: np.ndarray,
                   e2: Optional[np.ndarray], weights: Tuple[float, float]) -> np.ndarray:
    """idb:545-571."""
    e = e - neg_w * neg
    nrm = np.linalg.norm(e)
    if nrm > 0:
        return e / nrm
    return _restore(e1, e2, weights)


def apply_negatives(e: np.ndarray, negs: Sequence[np.ndarray], neg_ws: Sequence[float],
                    e1: np.ndarray, e2: Optional[np.ndarray],
                    weights: Tuple[float, float]) -> np.ndarray:
    """idb:573-604: subtraction is sequential, in list order; one normalise."""
    for v, w in zip(negs, neg_ws):
        e = e - w * v
    nrm = np.linalg.norm(e)
    if nrm > 0:
        return e / nrm
    return _restore(e1, e2, weights)


def compose_query(e1: np.ndarray, e2: Optional[np.ndarray] = None,
                  weights: Tuple[float, float] = (0.5, 0.5),
                  negatives: Sequence[np.ndarray] = (),
                  negative_weights: Sequence[float] = ()) -> np.ndarray:
    """The embedding ``search()`` hands to SQLite (idb:1340-1472), given the
    already-embedded inputs.  ``negatives`` is the reference's
    ``negative_embs_list`` (legacy single negative first, then the list)."""
    e1 = np.asarray(e1, dtype=np.float32)
    if e2 is not None:
        e2 = np.asarray(e2, dtype=np.float32)
        e, weights = blend_positive(e1, e2, weights)
    else:
        e = e1
    negs = [np.asarray(v, dtype=np.float32) for v in negatives]
    ws = list(negative_weights)
    if len(negs) == 1:
        e = apply_negative(e, negs[0], ws[0], e1, e2, weights)
    elif len(negs) > 1:
        e = apply_negatives(e, negs, ws, e1, e2, weights)
    return np.asarray(e, dtype=np.float32)
