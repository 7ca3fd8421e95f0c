#!/usr/bin/env python3
"""Generate tests/golden/reference_search.json by running the REFERENCE's own code.

Run in the authoring container only (needs /root/reference; the GPU box has no
copy).  What executes:

  * ``/root/reference/image_database.py`` is imported unmodified, with a stub
    ``sqlite_vec`` module injected first (the real extension is not installable
    here): ``sqlite_vec.load(conn)`` loads ``oracle/_build/vec_shim.so`` (the C
    restatement of ``vec_distance_cosine``) and ``serialize_float32`` is
    ``struct.pack``, as in sqlite-vec's published Python helper.
  * ``ImageDatabase`` is instantiated without ``__init__`` (no SigLIP weights
    offline); ``_get_text_embedding`` / ``_get_image_embedding`` are replaced by
    table look-ups of seeded vectors.
  * ``ImageDatabase.search()`` then runs as written — blend, negatives, guards,
    SQL text, ORDER BY/LIMIT in the real SQLite, similarity conversion, error
    envelope, duplicate filter — against a synthetic database in the reference
    schema whose ``vec0`` is a plain stand-in table.

So everything first-party to the reference is pinned by the reference itself;
only sqlite-vec's C arithmetic is the oracle's restatement (PARITY UNPINNED for
that part).  The inputs are regenerated from seeds by the tests; a SHA-256 of
each generated array is stored so a numpy stream change is detected, not
silently absorbed.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import struct
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
from clip_database_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_search.json")
DIM = 1152


def import_reference():
    stub = types.ModuleType("sqlite_vec")
    ref.build()

    def load(conn):
        conn.load_extension(ref.SHIM_PATH)

    def serialize_float32(vector):
        return struct.pack("%sf" % len(vector), *vector)

    stub.load = load
    stub.serialize_float32 = serialize_float32
    sys.modules["sqlite_vec"] = stub
    sys.path.insert(0, "/root/reference")
    with contextlib.redirect_stdout(io.StringIO()):
        import image_database
    return image_database


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_cases():
    """Each case: how to make the database and the call; see tests/test_golden.py
    (``materialise_case``) for the matching reader."""
    cases = []

    def add(name, **kw):
        kw["name"] = name
        cases.append(kw)

    base = dict(n=6000, rows_seed=1234, query_seed=99)
    add("single_k20", **base, k=20, show_duplicates=True)
    add("single_k1", **base, k=1, show_duplicates=True)
    add("single_k100", **base, k=100, show_duplicates=True)
    add("self_match_ties", **base, k=10, show_duplicates=True, dup_rows=[[7, 900], [7, 1500], [7, 4000]],
        query_row=7)
    add("blend_07_03_negative", **base, k=20, show_duplicates=True, query2_seed=100,
        weights=[0.7, 0.3], negative_seeds=[101], negative_weights=[0.5], legacy_negative=True)
    add("blend_unnormalised_weights", **base, k=20, show_duplicates=True, query2_seed=100, weights=[2.0, 6.0])
    add("blend_zero_weights", **base, k=20, show_duplicates=True, query2_seed=100, weights=[0.0, 0.0])
    add("three_negatives", **base, k=20, show_duplicates=True, negative_seeds=[101, 102, 103],
        negative_weights=[0.5, 0.25, 1.5])
    add("legacy_plus_list_negatives", **base, k=20, show_duplicates=True, negative_seeds=[101, 102],
        negative_weights=[0.3, 0.6], legacy_negative=True)
    add("blend_cancels_to_zero", **base, k=20, show_duplicates=True, query2_negated=True, weights=[0.5, 0.5])
    add("negative_cancels_to_zero", **base, k=20, show_duplicates=True, negative_is_query=True,
        negative_weights=[1.0], legacy_negative=True)
    # sparse power-of-two vectors: the blend and its norm are exact in any summation order, so
    # "blend minus 1.0 x blend" is exactly zero on every machine and the re-blend branch fires
    add("negative_cancels_blend_restored", **base, k=20, show_duplicates=True, sparse_pair=True,
        weights=[0.5, 0.5], negative_is_blend=True, negative_weights=[1.0], legacy_negative=True)
    add("folder_filter", **base, k=20, show_duplicates=True, filter_folders=["/data/photos/b"])
    add("folder_filter_two_case_insensitive", **base, k=20, show_duplicates=True,
        filter_folders=["/DATA/photos/a", "/data/scans/"])
    add("folder_filter_wildcard_chars", **base, k=20, show_duplicates=True, special_paths=True,
        filter_folders=["/data/100%_done"])
    add("duplicate_filter_default", **base, k=20, show_duplicates=False, near_dup_of_top=True)
    add("fp16_normalised_rows", **base, k=20, show_duplicates=True, fp16=True)
    add("orphan_rows", **base, k=20, show_duplicates=True, drop_mapping_top=2, drop_image_top=1)
    add("zero_row_gives_empty", **base, k=20, show_duplicates=True, zero_rows=[5])
    add("k_zero", n=300, rows_seed=7, query_seed=99, k=0, show_duplicates=True)
    add("k_exceeds_rows", n=300, rows_seed=7, query_seed=99, k=350, show_duplicates=True)
    add("k_negative_unlimited", n=300, rows_seed=7, query_seed=99, k=-1, show_duplicates=True)
    return cases


def materialise_case(case):
    """(rows, paths, call kwargs, named vectors) for a case — shared with the tests."""
    n = case["n"]
    rows = synth.unit_rows(n, DIM, case["rows_seed"])
    q = synth.unit_rows(1, DIM, case["query_seed"])[0]
    if case.get("fp16"):
        rows = synth.fp16_normalised(rows)
        q = synth.fp16_normalised(q)
    for dst_src in case.get("dup_rows", []):
        rows[dst_src[1]] = rows[dst_src[0]]
    if "query_row" in case:
        q = rows[case["query_row"]].copy()
    for z in case.get("zero_rows", []):
        rows[z] = 0
    paths = synth.default_paths(n)
    if case.get("special_paths"):
        for i in range(0, n, 5):
            paths[i] = f"/data/100%_done/img_{i:08d}.jpg"
        for i in range(1, n, 5):
            paths[i] = f"/data/100x_done/img_{i:08d}.jpg"     # must NOT match the escaped pattern
    vectors = {"q1": q}
    kwargs = dict(k=case["k"], show_duplicates=case["show_duplicates"])
    if case.get("sparse_pair"):
        q = np.zeros(DIM, dtype=np.float32)
        q[0:4] = 0.5
        q2 = np.zeros(DIM, dtype=np.float32)
        q2[10:14] = 0.5
        vectors["q1"] = q
        vectors["q2"] = q2
    if "query2_seed" in case:
        vectors["q2"] = synth.unit_rows(1, DIM, case["query2_seed"])[0]
    if case.get("query2_negated"):
        vectors["q2"] = -q
    if "q2" in vectors:
        kwargs["query2"] = "q2"
    if "weights" in case:
        kwargs["weights"] = tuple(case["weights"])
    negs = []
    for s in case.get("negative_seeds", []):
        negs.append(synth.unit_rows(1, DIM, s)[0])
    if case.get("negative_is_query"):
        negs.append(q.copy())
    if case.get("negative_is_blend"):
        # only used with sparse_pair: 8 entries of 0.25, norm sqrt(0.5), all exact
        b = np.float32(0.5) * q + np.float32(0.5) * vectors["q2"]
        negs.append(b / np.sqrt(np.float32(0.5)))
    for i, v in enumerate(negs):
        vectors[f"n{i}"] = v
    nws = list(case.get("negative_weights", []))
    if negs:
        if case.get("legacy_negative"):
            kwargs["negative_query"] = "n0"
            kwargs["negative_weight"] = nws[0]
            if len(negs) > 1:
                kwargs["negative_queries"] = [f"n{i}" for i in range(1, len(negs))]
                kwargs["negative_weights"] = nws[1:]
        else:
            kwargs["negative_queries"] = [f"n{i}" for i in range(len(negs))]
            kwargs["negative_weights"] = nws
    if "filter_folders" in case:
        kwargs["filter_folders"] = list(case["filter_folders"])
    return rows, paths, kwargs, vectors


def post_adjust(case, rows, q):
    """Adjustments that need a first look at the ranking (near duplicates of the
    best hit, orphaning the best hits).  Returns (rows, drop_mapping, drop_image)."""
    drop_m, drop_i = [], []
    if case.get("near_dup_of_top") or case.get("drop_mapping_top") or case.get("drop_image_top"):
        _, _, seq, _ = ref.knn(rows, q, 5)
        if case.get("near_dup_of_top"):
            # rows 3000/3001/3002 become (near) copies of the best hit: identical sign code,
            # one differing sign, three differing signs (the last is NOT a duplicate at tolerance 2)
            top = rows[seq[0]].copy()
            small = np.argsort(np.abs(top))[:3]
            rows[3000] = top
            v = top.copy(); v[small[0]] = -v[small[0]]; rows[3001] = v
            v = top.copy(); v[small] = -v[small]; rows[3002] = v
        drop_m = [int(s) for s in seq[:case.get("drop_mapping_top", 0)]]
        drop_i = [int(s) for s in seq[2:2 + case.get("drop_image_top", 0)]]
    return rows, drop_m, drop_i


def main():
    idb = import_reference()
    tmp = tempfile.mkdtemp(prefix="golden_")
    out = {"generated_by": "reference image_database.py ImageDatabase.search(), sqlite_vec stubbed "
                           "(load -> oracle vec_shim.so, vec0 -> plain stand-in table)",
           "numpy": np.__version__, "dim": DIM, "cases": []}
    for case in build_cases():
        rows, paths, kwargs, vectors = materialise_case(case)
        rows, drop_m, drop_i = post_adjust(case, rows, vectors["q1"])
        db_path = os.path.join(tmp, case["name"] + ".db")
        synth.write_reference_db(db_path, rows, paths, drop_mapping_for=drop_m, drop_image_for=drop_i)
        db = object.__new__(idb.ImageDatabase)
        db.db_path = db_path
        db.embedding_dim = DIM
        db._get_text_embedding = lambda text, _v=vectors: _v[text]
        db._get_image_embedding = lambda path, _v=vectors: _v[path]
        with contextlib.redirect_stdout(io.StringIO()):
            results = db.search("q1", **kwargs)
        pos = {p: i for i, p in enumerate(paths)}
        rec = dict(case)
        rec["rows_sha256"] = sha(rows)
        rec["drop_mapping_for"] = drop_m
        rec["drop_image_for"] = drop_i
        rec["expected_positions"] = [pos[p] for p, _ in results]
        rec["expected_similarities"] = [float(s) for _, s in results]
        out["cases"].append(rec)
        print(f"{case['name']:40s} -> {len(results)} results, first {results[:1]}")
        os.remove(db_path)

    # blend-only vectors straight from the reference's methods
    blends = []
    db = object.__new__(idb.ImageDatabase)
    rng = np.random.default_rng(5)
    e1 = synth.unit_rows(1, DIM, 201)[0]
    e2 = synth.unit_rows(1, DIM, 202)[0]
    n1 = synth.unit_rows(1, DIM, 203)[0]
    n2 = synth.unit_rows(1, DIM, 204)[0]
    with contextlib.redirect_stdout(io.StringIO()):
        one = db._apply_negative_embedding(e1, n1, 0.5, e1, None, (0.5, 0.5))
        many = db._apply_multiple_negative_embeddings(e1, [n1, n2], [0.5, 0.8], e1, None, (0.5, 0.5))
        zero1 = db._apply_negative_embedding(e1, e1, 1.0, e1, None, (0.5, 0.5))
        zero2 = db._apply_negative_embedding(e1, e1, 1.0, e1, e2, (0.7, 0.3))
    del rng
    for name, v in (("one_negative", one), ("two_negatives", many), ("zero_restores_e1", zero1),
                    ("zero_reblends", zero2)):
        blends.append({"name": name, "seeds": [201, 202, 203, 204],
                       "out": [float(x) for x in np.asarray(v, dtype=np.float32)]})
    out["blend_cases"] = blends
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
