"""Rebuild the inputs of a golden case (shared by the CPU and GPU golden tests)."""
import hashlib
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(_HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


def inputs_for(case):
    """(rows, paths, search kwargs, named vectors, drop_mapping, drop_image), verified
    against the SHA-256 recorded when the reference produced the expected output."""
    rows, paths, kwargs, vectors = make_golden.materialise_case(case)
    rows, drop_m, drop_i = make_golden.post_adjust(case, rows, vectors["q1"])
    digest = hashlib.sha256(np.ascontiguousarray(rows).tobytes()).hexdigest()
    assert digest == case["rows_sha256"], "seeded inputs differ from the ones the golden file was made with"
    assert drop_m == case["drop_mapping_for"] and drop_i == case["drop_image_for"]
    return rows, paths, kwargs, vectors, drop_m, drop_i


def embedding_call(kwargs, vectors):
    """Translate the reference-style search kwargs (names of vectors) into the
    arguments of ``search_embedding`` / the oracle's ``compose_query``."""
    e1 = vectors["q1"]
    e2 = vectors[kwargs["query2"]] if "query2" in kwargs else None
    weights = kwargs.get("weights", (0.5, 0.5))
    negs, ws = [], []
    if "negative_query" in kwargs:
        negs.append(vectors[kwargs["negative_query"]])
        ws.append(kwargs.get("negative_weight", 0.5))
    for i, name in enumerate(kwargs.get("negative_queries", [])):
        negs.append(vectors[name])
        nw = kwargs.get("negative_weights")
        ws.append(nw[i] if nw and i < len(nw) else kwargs.get("negative_weight", 0.5))
    return e1, e2, weights, negs, ws
