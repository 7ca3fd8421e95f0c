"""A resident copy next to the reference's own writer: tests/golden/reference_rescan.json holds the row-level writes
``ImageDatabase._commit_batch`` (image_database.py:1098-1204) made when files were (re)scanned, and what the reference's
``search()`` returned after each stage.  Replayed here with plain SQL: the real SQLite statement must reproduce the
recorded results (CPU: pins the replay), and ``ImageDatabase.refresh()`` on the GPU must follow them."""
import sqlite3

import numpy as np
import pytest

from clip_database_b200 import synth
from oracle import sql_harness

import golden_cases
from conftest import have_gpu, tol

DIM = 1152


def fresh_database(tmp_path, g):
    rows = synth.unit_rows(g["n"], DIM, g["rows_seed"])
    db_path = str(tmp_path / "rescan.db")
    synth.write_reference_db(db_path, rows, synth.default_paths(g["n"]))
    return db_path, synth.unit_rows(1, DIM, g["query_seed"])[0]


def assert_matches(got, paths, sims):
    assert [p for p, _ in got] == paths
    w = np.asarray(sims, dtype=np.float64)
    assert np.all(np.abs(np.array([s for _, s in got]) - w) <= tol(1.0 - w))


def test_replayed_writes_reproduce_the_reference_results_in_sqlite(tmp_path):
    g = golden_cases.rescan_golden()
    db_path, q = fresh_database(tmp_path, g)
    assert_matches(sql_harness.reference_search(db_path, q, g["k"]), g["initial"]["paths"], g["initial"]["similarities"])
    conn = sqlite3.connect(db_path)
    for stage in g["stages"]:
        golden_cases.apply_rescan_stage(conn, stage)
        assert_matches(sql_harness.reference_search(db_path, q, g["k"]), stage["paths"], stage["similarities"])
    conn.close()
    # what the writer does to a modified file: the image row is re-keyed, its old vec0 row stays behind, orphaned
    first = g["stages"][0]
    assert len(first["images"]["deleted"]) == 1 and first["vec0"]["deleted"] == []
    assert g["initial"]["paths"][0] not in first["paths"]


@pytest.mark.gpu
@pytest.mark.parametrize("batch_store", [False, True], ids=["f32scan", "bf16preselect"])
def test_refresh_follows_the_references_own_writer(tmp_path, batch_store):
    assert have_gpu(), "GPU tests selected but no CUDA device is visible"
    from clip_database_b200 import ImageDatabase
    g = golden_cases.rescan_golden()
    db_path, q = fresh_database(tmp_path, g)
    db = ImageDatabase(db_path, device=0, batch_store=batch_store)
    conn = sqlite3.connect(db_path)
    try:
        assert_matches(db.search_embedding(q, k=g["k"], show_duplicates=True), g["initial"]["paths"],
                       g["initial"]["similarities"])
        for stage in g["stages"]:
            golden_cases.apply_rescan_stage(conn, stage)
            changed = db.refresh()
            assert changed >= len(stage["vec0"]["upserted"]), stage["name"]
            assert db.reloads == 1, "appends and re-keyed files must not need a reload"
            assert_matches(db.search_embedding(q, k=g["k"], show_duplicates=True), stage["paths"], stage["similarities"])
        assert db.refresh() == 0
    finally:
        conn.close()
        db.close()
