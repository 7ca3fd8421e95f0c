"""The oracle's three tiers agree with each other and with SQLite's own semantics."""
import os

import numpy as np
import pytest

from clip_database_b200 import synth
from oracle import blend as oblend
from oracle import ref, sql_harness

from conftest import tol

DIM = 1152


@pytest.fixture(scope="module")
def small():
    rows = synth.unit_rows(4000, DIM, 1234)
    rows[900] = rows[7]
    rows[1500] = rows[7]
    queries = synth.unit_rows(6, DIM, 99)
    return rows, queries


@pytest.fixture(scope="module")
def small_db(small, tmp_path_factory):
    rows, _ = small
    path = str(tmp_path_factory.mktemp("oracle") / "small.db")
    synth.write_reference_db(path, rows)
    return path


def test_ref_matches_exact_within_tolerance(small):
    rows, queries = small
    for q in queries:
        d32 = ref.distances(rows, q)
        d64 = ref.distances_f64(rows, q)
        assert np.all(np.abs(d32 - d64) <= 0.1 * tol(d64))      # measured ~2e-7, bound 1e-5
        ids, d, seq, n_nan = ref.knn(rows, q, 20)
        eids, ed, eseq = ref.exact_knn(rows, q, 20)
        assert n_nan == 0
        assert np.array_equal(seq, eseq)
        assert np.all(np.abs(d - ed) <= tol(ed))


def test_ref_matches_sqlite_statement(small, small_db):
    """oracle_ref == the reference's SQL run by the real SQLite (ids, order, distances)."""
    rows, queries = small
    conn, provider = sql_harness.connect(small_db)
    assert "shim" in provider or provider == "sqlite-vec" or "callback" in provider
    for q in queries[:3]:
        for k in (1, 20, 100):
            got = sql_harness.run_statement(conn, q, k, with_rowid=True)
            ids, d, seq, _ = ref.knn(rows, q, k, rowids=np.arange(1, rows.shape[0] + 1))
            assert [r[1] for r in got] == ids.tolist()
            assert np.array_equal(np.array([r[2] for r in got], dtype=np.float32), d)
    conn.close()


def test_exact_ties_come_back_in_rowid_order(small, small_db):
    rows, _ = small
    ids, d, seq, _ = ref.knn(rows, rows[7], 5)
    assert seq[:3].tolist() == [7, 900, 1500]
    assert d[0] == d[1] == d[2]
    conn, _ = sql_harness.connect(small_db)
    got = sql_harness.run_statement(conn, rows[7], 3, with_rowid=True)
    assert [r[1] for r in got] == [8, 901, 1501]
    # a later row that ties the current k-th is not admitted
    got2 = sql_harness.run_statement(conn, rows[7], 2, with_rowid=True)
    assert [r[1] for r in got2] == [8, 901]
    conn.close()


@pytest.mark.parametrize("k", [0, 1, 20, 3999, 4000, 4050, -1])
def test_k_edges(small, small_db, k):
    rows, queries = small
    ids, d, seq, _ = ref.knn(rows, queries[0], k)
    expect = rows.shape[0] if (k < 0 or k > rows.shape[0]) else k
    assert len(ids) == expect
    conn, _ = sql_harness.connect(small_db)
    got = sql_harness.run_statement(conn, queries[0], k, with_rowid=True)
    assert [r[1] - 1 for r in got] == seq.tolist()
    conn.close()


def test_zero_row_is_null_and_sorts_first(tmp_path):
    rows = synth.unit_rows(200, DIM, 5)
    rows[17] = 0
    q = synth.unit_rows(1, DIM, 6)[0]
    ids, d, seq, n_nan = ref.knn(rows, q, 10)
    assert n_nan == 1 and 17 not in seq
    db = str(tmp_path / "z.db")
    synth.write_reference_db(db, rows)
    conn, _ = sql_harness.connect(db)
    got = sql_harness.run_statement(conn, q, 3, with_rowid=True)
    assert got[0][1] == 18 and got[0][2] is None           # NaN -> NULL, NULLs first
    conn.close()
    assert sql_harness.reference_search(db, q, 3) == []   # `1.0 - None` inside the envelope


def test_mask_restates_where_clause(small, small_db):
    rows, queries = small
    paths = synth.default_paths(rows.shape[0])
    mask = np.array([p.startswith("/data/photos/b/") for p in paths], dtype=np.uint8)
    ids, d, seq, _ = ref.knn(rows, queries[1], 20, mask=mask)
    conn, _ = sql_harness.connect(small_db)
    got = sql_harness.run_statement(conn, queries[1], 20, ["/data/photos/b"], with_rowid=True)
    assert [r[1] - 1 for r in got] == seq.tolist()
    conn.close()


def test_multithreaded_distances_identical(small):
    rows, queries = small
    assert np.array_equal(ref.distances(rows, queries[0]), ref.distances(rows, queries[0], threads=4))


def test_fp16_normalised_vectors_need_true_cosine():
    """||q|| ~ 0.9999 after fp16 normalisation: a bare dot product
    would miss the 1e-5 tolerance, the cosine restatement does not."""
    rows = synth.fp16_normalised(synth.unit_rows(500, DIM, 3))
    q = synth.fp16_normalised(synth.unit_rows(1, DIM, 4)[0])
    d = ref.distances(rows, q)
    bare = 1.0 - rows.astype(np.float64) @ q.astype(np.float64)
    assert np.max(np.abs(bare - d)) > 1e-6           # the norms matter...
    d64 = ref.distances_f64(rows, q)
    assert np.all(np.abs(d - d64) <= tol(d64))       # ...and the restatement handles them


def test_blend_oracle_weight_rules():
    e1 = synth.unit_rows(1, DIM, 11)[0]
    e2 = synth.unit_rows(1, DIM, 12)[0]
    a = oblend.compose_query(e1, e2, (2.0, 6.0))
    b = oblend.compose_query(e1, e2, (0.25, 0.75))
    assert np.array_equal(a, b)
    z = oblend.compose_query(e1, e2, (0.0, 0.0))
    h = oblend.compose_query(e1, e2, (0.5, 0.5))
    assert np.array_equal(z, h)
    assert np.array_equal(oblend.compose_query(e1, -e1, (0.5, 0.5)), e1)      # zero norm -> e1
    assert np.array_equal(oblend.compose_query(e1, None, negatives=[e1], negative_weights=[1.0]), e1)
    assert abs(np.linalg.norm(oblend.compose_query(e1, e2, (0.7, 0.3), [e2], [0.5])) - 1.0) < 1e-6
