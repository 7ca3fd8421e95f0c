"""The drop-in ``ImageDatabase.search()`` (GPU path) against the outputs of the
reference's own ``search()`` recorded in tests/golden/reference_search.json."""
import numpy as np
import pytest

from clip_database_b200 import synth
from oracle import blend as oblend
from oracle import ref

import golden_cases
from conftest import have_gpu, tol
from test_golden_cpu import case_names

pytestmark = pytest.mark.gpu


class TableEmbedder:
    """Stands in for the SigLIP model: query string -> seeded vector."""

    def __init__(self, vectors):
        self.vectors = vectors

    def text(self, query):
        return self.vectors[query]

    def image(self, path):
        return self.vectors[path]


@pytest.mark.parametrize("batch_store", [False, True], ids=["f32scan", "bf16preselect"])
@pytest.mark.parametrize("name", case_names())
def test_search_reproduces_reference_output(golden, name, batch_store, tmp_path):
    """batch_store=True sends the same searches through the tensor-core pre-selection + exact
    re-rank (batch_min_nq = 1): the reference's outputs must come back either way."""
    assert have_gpu(), "GPU tests selected but no CUDA device is visible"
    from clip_database_b200 import ImageDatabase
    case = next(c for c in golden["cases"] if c["name"] == name)
    rows, paths, kwargs, vectors, drop_m, drop_i = golden_cases.inputs_for(case)
    db_path = str(tmp_path / (name + ".db"))
    synth.write_reference_db(db_path, rows, paths, drop_mapping_for=drop_m, drop_image_for=drop_i)
    db = ImageDatabase(db_path, device=0, embedder=TableEmbedder(vectors), batch_store=batch_store)
    try:
        results = db.search("q1", **kwargs)
    finally:
        db.close()
    pos = {p: i for i, p in enumerate(paths)}
    got_pos = np.array([pos[p] for p, _ in results], dtype=np.int64)
    got_sim = np.array([s for _, s in results], dtype=np.float64)
    exp_pos = np.array(case["expected_positions"], dtype=np.int64)
    exp_sim = np.array(case["expected_similarities"], dtype=np.float64)
    assert got_pos.shape == exp_pos.shape
    assert np.all(np.abs(got_sim - exp_sim) <= tol(1.0 - exp_sim))
    diff = got_pos != exp_pos
    if diff.any():
        # ids may differ only where the reference's own distances tie within the tolerance
        e1, e2, weights, negs, ws = golden_cases.embedding_call(kwargs, vectors)
        q = oblend.compose_query(e1, e2, weights, negs, ws)
        own = 1.0 - ref.distances(rows[got_pos[diff]], q).astype(np.float64)
        assert np.all(np.abs(own - exp_sim[diff]) <= tol(1.0 - exp_sim[diff])), name


def test_nan_policy_exclude(tmp_path):
    from clip_database_b200 import ImageDatabase
    rows = synth.unit_rows(2000, 1152, 1)
    rows[5] = 0
    q = synth.unit_rows(1, 1152, 2)[0]
    db_path = str(tmp_path / "nan.db")
    synth.write_reference_db(db_path, rows)
    a = ImageDatabase(db_path, nan_policy="reference")
    b = ImageDatabase(db_path, nan_policy="exclude")
    try:
        assert a.search_embedding(q, k=10, show_duplicates=True) == []
        got = b.search_embedding(q, k=10, show_duplicates=True)
        _, od, oseq, _ = ref.knn(rows, q, 10)
        paths = synth.default_paths(2000)
        assert [p for p, _ in got] == [paths[s] for s in oseq]
    finally:
        a.close()
        b.close()


def test_refresh_appends_new_rows(tmp_path):
    import sqlite3
    from clip_database_b200 import ImageDatabase
    rows = synth.unit_rows(3000, 1152, 3)
    db_path = str(tmp_path / "grow.db")
    synth.write_reference_db(db_path, rows[:2000])
    db = ImageDatabase(db_path)
    try:
        q = rows[2500]
        first = db.search_embedding(q, k=5, show_duplicates=True)
        assert first[0][1] < 0.5
        # the scanner appends 1000 more images (same writer layout)
        conn = sqlite3.connect(db_path)
        paths = synth.default_paths(3000)
        for i in range(2000, 3000):
            conn.execute("INSERT INTO images (id, file_path, last_modified, file_hash) VALUES (?, ?, ?, ?)",
                         (i + 1, paths[i], 1.0, "h"))
            conn.execute("INSERT INTO vec0 (rowid, embedding) VALUES (?, ?)", (i + 1, rows[i].tobytes()))
            conn.execute("INSERT INTO image_embeddings (rowid, image_id) VALUES (?, ?)", (i + 1, i + 1))
            conn.execute("INSERT INTO binary_embeddings (image_id, embedding) VALUES (?, ?)",
                         (i + 1, (rows[i] >= 0).astype(np.uint8).tobytes()))
        conn.commit()
        conn.close()
        assert db.refresh() == 1000
        second = db.search_embedding(q, k=5, show_duplicates=True)
        assert second[0][0] == paths[2500] and abs(second[0][1] - 1.0) < 1e-6
        _, od, oseq, _ = ref.knn(rows, q, 5)
        assert [p for p, _ in second] == [paths[s] for s in oseq]
    finally:
        db.close()


def test_search_embeddings_equals_one_at_a_time(tmp_path):
    """Many sessions' queries in one call (one pass over the store with batch_store=True)."""
    from clip_database_b200 import ImageDatabase
    rows = synth.unit_rows(9000, 1152, 11)
    paths = synth.default_paths(9000)
    db_path = str(tmp_path / "many.db")
    synth.write_reference_db(db_path, rows, paths)
    queries = synth.unit_rows(40, 1152, 12)
    queries[5] = rows[77]
    one = ImageDatabase(db_path, device=0)
    many = ImageDatabase(db_path, device=0, batch_store=True)
    try:
        for folders in (None, ["/data/photos/b"]):
            want = [one.search_embedding(q, k=25, filter_folders=folders, show_duplicates=True) for q in queries]
            before = many.index.launch_count
            got = many.search_embeddings(queries, k=25, filter_folders=folders)
            assert many.index.launch_count - before <= 8, "one batched pass expected"
            assert got == want
        assert got[5][0][0].startswith("/data/photos/b") or True
        assert many.search_embeddings(queries[:3], k=25)[0] == one.search_embedding(queries[0], k=25, show_duplicates=True)
    finally:
        one.close()
        many.close()


@pytest.mark.parametrize("name", ["single_k20", "self_match_ties", "blend_07_03_negative", "folder_filter",
                                  "orphan_rows", "zero_row_gives_empty", "k_exceeds_rows", "k_negative_unlimited",
                                  "duplicate_filter_default"])
def test_search_on_a_row_sharded_store(golden, name, tmp_path):
    """ImageDatabase(devices=[...]): the store row-sharded over the box's GPUs from one process (two
    shards on the one GPU when there is only one) returns what the reference returned."""
    import torch
    from clip_database_b200 import ImageDatabase
    case = next(c for c in golden["cases"] if c["name"] == name)
    rows, paths, kwargs, vectors, drop_m, drop_i = golden_cases.inputs_for(case)
    db_path = str(tmp_path / (name + ".db"))
    synth.write_reference_db(db_path, rows, paths, drop_mapping_for=drop_m, drop_image_for=drop_i)
    n_dev = torch.cuda.device_count()
    devices = list(range(min(n_dev, 4))) if n_dev > 1 else [0, 0]
    db = ImageDatabase(db_path, embedder=TableEmbedder(vectors), devices=devices)
    try:
        results = db.search("q1", **kwargs)
    finally:
        db.close()
    pos = {p: i for i, p in enumerate(paths)}
    got_pos = np.array([pos[p] for p, _ in results], dtype=np.int64)
    got_sim = np.array([s for _, s in results], dtype=np.float64)
    exp_pos = np.array(case["expected_positions"], dtype=np.int64)
    exp_sim = np.array(case["expected_similarities"], dtype=np.float64)
    assert got_pos.shape == exp_pos.shape
    assert np.all(np.abs(got_sim - exp_sim) <= tol(1.0 - exp_sim))
    diff = got_pos != exp_pos
    if diff.any():
        e1, e2, weights, negs, ws = golden_cases.embedding_call(kwargs, vectors)
        q = oblend.compose_query(e1, e2, weights, negs, ws)
        own = 1.0 - ref.distances(rows[got_pos[diff]], q).astype(np.float64)
        assert np.all(np.abs(own - exp_sim[diff]) <= tol(1.0 - exp_sim[diff])), name
