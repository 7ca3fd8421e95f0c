"""Sign-code fallback: the oracle reproduces what the REFERENCE's own
search() returned on binary-only databases (tests/golden/reference_binary_search.json, made by
make_golden_binary.py), and the host-side pieces of the CUDA path (scan order, tie sequence,
Python-slice k) agree with the real SQLite."""
import json
import os

import numpy as np
import pytest

from clip_database_b200 import database, loader, synth
from oracle import binary as obinary
from oracle import blend as oblend
from oracle import sql_harness

import golden_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_binary_search.json")


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


def case_names():
    return [c["name"] for c in load_golden()["cases"]]


def inputs_for(case):
    import hashlib
    mg = golden_cases.make_golden
    rows, _, kwargs, vectors = mg.materialise_case(case)
    assert hashlib.sha256(np.ascontiguousarray(rows).tobytes()).hexdigest() == case["rows_sha256"]
    n = case["n"]
    if case.get("scrambled"):
        folders = ("/data/photos/a", "/data/photos/b", "/data/scans")
        paths = [f"{folders[i % 3]}/img_{(i * 7919) % 100003:08d}.jpg" for i in range(n)]
    else:
        paths = synth.default_paths(n)
    return rows, paths, kwargs, vectors


@pytest.mark.parametrize("name", case_names())
def test_oracle_reproduces_reference_binary_search(name, tmp_path):
    case = next(c for c in load_golden()["cases"] if c["name"] == name)
    rows, paths, kwargs, vectors = inputs_for(case)
    e1, e2, weights, negs, ws = golden_cases.embedding_call(kwargs, vectors)
    q = oblend.compose_query(e1, e2, weights, negs, ws)
    db = str(tmp_path / "b.db")
    synth.write_reference_db(db, rows, paths, vectors=False)

    # literal restatement on the real SQLite
    results = sql_harness.reference_binary_search(db, q, kwargs["k"], kwargs.get("filter_folders"))
    if not kwargs["show_duplicates"] and results:
        results = sql_harness.reference_filter_duplicates(db, results, 2)
    pos = {p: i for i, p in enumerate(paths)}
    assert [pos[p] for p, _ in results] == case["expected_positions"]
    assert [s for _, s in results] == case["expected_similarities"]

    # vectorised oracle + the host logic the CUDA path uses (scan order from the loader, tie
    # sequence = file_path rank when SQLite walks the path index, Python-slice k)
    host = loader.read_codes(db)
    assert host.file_paths == paths and np.array_equal(host.codes, (rows >= 0).astype(np.uint8))
    order = None
    k = kwargs["k"]
    if "filter_folders" in kwargs:
        admitted = database.like_prefix_mask(paths, kwargs["filter_folders"])
        by_path = sorted(range(len(paths)), key=lambda i: paths[i].encode())
        order = [i for i in by_path if admitted[i]]
    opos, oscore = obinary.search(host.codes, obinary.sign_code(q), k, wrap=True, order=order)
    if kwargs["show_duplicates"]:
        assert opos.tolist() == case["expected_positions"]
        assert [float(s) / 1152 for s in oscore] == case["expected_similarities"]


def test_filtered_statement_scan_order_is_what_the_host_assumes(tmp_path):
    """The CUDA path's tie sequence for filtered searches mirrors the plan THIS SQLite picks."""
    rows = synth.unit_rows(400, 1152, 3)
    folders = ("/data/photos/a", "/data/photos/b")
    paths = [f"{folders[i % 2]}/img_{(i * 7919) % 1009:06d}.jpg" for i in range(400)]
    db = str(tmp_path / "o.db")
    synth.write_reference_db(db, rows, paths, vectors=False)
    import sqlite3
    conn = sqlite3.connect(db)
    where, params = sql_harness.where_clause_and_params(["/data/photos/b"])
    got = [r[2] for r in conn.execute(sql_harness.BINARY_SQL.format(where_clause=where), params)]
    plan = conn.execute("EXPLAIN QUERY PLAN " + sql_harness.BINARY_SQL.format(where_clause=where), params).fetchall()
    unfiltered = [r[2] for r in conn.execute(sql_harness.BINARY_SQL.format(where_clause=""))]
    conn.close()
    assert unfiltered == paths                                       # binary_embeddings rowid order
    admitted = [p for p in paths if p.startswith("/data/photos/b/")]
    if str(plan[0][-1]).startswith("SCAN i"):
        assert got == sorted(admitted, key=lambda p: p.encode())     # file_path index order
    else:
        assert got == admitted


def test_oracle_scores_wrap_like_numpy_uint8():
    rng = np.random.default_rng(1)
    codes = (rng.random((64, 1152)) < 0.8).astype(np.uint8)
    q = np.ones(1152, dtype=np.uint8)
    lit = np.array([int(np.dot(q, c)) for c in codes])               # numpy's own uint8 result
    assert np.array_equal(obinary.scores(codes, q, wrap=True), lit)
    assert np.array_equal(obinary.scores(codes, q, wrap=False), codes.sum(axis=1))
    assert (obinary.scores(codes, q, wrap=False) > 255).all()
