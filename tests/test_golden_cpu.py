"""The oracle stack reproduces what the REFERENCE's own search() returned.

tests/golden/reference_search.json was produced by running
/root/reference/image_database.py ``ImageDatabase.search()`` (see make_golden.py).
Here the oracle (numpy blend restatement + oracle_ref.c KNN + host mask / duplicate
logic) is checked against every one of those outputs; the GPU tests check the CUDA
path against the same file.
"""
import numpy as np
import pytest

from clip_database_b200 import database, synth
from oracle import blend as oblend
from oracle import ref

import golden_cases
from conftest import tol


def case_names(golden_path="tests/golden/reference_search.json"):
    import json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, golden_path)) as f:
        return [c["name"] for c in json.load(f)["cases"]]


@pytest.mark.parametrize("name", case_names())
def test_oracle_reproduces_reference_search(golden, name):
    case = next(c for c in golden["cases"] if c["name"] == name)
    rows, paths, kwargs, vectors, drop_m, drop_i = golden_cases.inputs_for(case)
    e1, e2, weights, negs, ws = golden_cases.embedding_call(kwargs, vectors)
    q = oblend.compose_query(e1, e2, weights, negs, ws)

    admitted = np.ones(rows.shape[0], dtype=bool)
    admitted[drop_m] = False
    admitted[drop_i] = False
    if "filter_folders" in kwargs:
        admitted &= database.like_prefix_mask(paths, kwargs["filter_folders"])
    ids, d, seq, n_nan = ref.knn(rows, q, kwargs["k"], mask=admitted.astype(np.uint8))
    if n_nan > 0:
        results = []          # NULL sorts first and `1.0 - None` raises inside the envelope
    else:
        results = [(paths[s], 1.0 - float(x)) for s, x in zip(seq, d)]
    if not kwargs["show_duplicates"] and results:
        codes = {paths[s]: (rows[s] >= 0).astype(np.uint8) for s in seq}
        results = database.filter_duplicates(results, codes, 2)

    pos = {p: i for i, p in enumerate(paths)}
    got_pos = [pos[p] for p, _ in results]
    assert got_pos == case["expected_positions"]
    exp_sim = np.array(case["expected_similarities"])
    got_sim = np.array([s for _, s in results])
    # the blended query may differ in the last bit from the reference's (BLAS summation
    # order of the norm), so similarities carry the float tolerance, not bit equality
    assert np.all(np.abs(got_sim - exp_sim) <= tol(1.0 - exp_sim))


def test_blend_oracle_matches_reference_methods(golden):
    by = {b["name"]: np.array(b["out"], dtype=np.float32) for b in golden["blend_cases"]}
    e1 = synth.unit_rows(1, 1152, 201)[0]
    e2 = synth.unit_rows(1, 1152, 202)[0]
    n1 = synth.unit_rows(1, 1152, 203)[0]
    n2 = synth.unit_rows(1, 1152, 204)[0]
    assert np.allclose(oblend.apply_negative(e1, n1, 0.5, e1, None, (0.5, 0.5)), by["one_negative"],
                       rtol=1e-6, atol=1e-8)
    assert np.allclose(oblend.apply_negatives(e1, [n1, n2], [0.5, 0.8], e1, None, (0.5, 0.5)),
                       by["two_negatives"], rtol=1e-6, atol=1e-8)
    assert np.array_equal(oblend.apply_negative(e1, e1, 1.0, e1, None, (0.5, 0.5)), by["zero_restores_e1"])
    assert np.allclose(oblend.apply_negative(e1, e1, 1.0, e1, e2, (0.7, 0.3)), by["zero_reblends"],
                       rtol=1e-6, atol=1e-8)
