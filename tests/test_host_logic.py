"""Host-side pieces of the drop-in: loader, folder mask, duplicate filter."""
import os
import sqlite3

import numpy as np
import pytest

from clip_database_b200 import database, loader, synth
from oracle import sql_harness

DIM = 1152


def test_loader_plain_and_shadow_agree(tmp_path):
    rows = synth.unit_rows(2500, DIM, 21)
    a, b = str(tmp_path / "a.db"), str(tmp_path / "b.db")
    synth.write_reference_db(a, rows, rowid_start=5)
    synth.write_reference_db(b, rows, rowid_start=5, vec0_layout="shadow")
    ha, hb = loader.read_store(a), loader.read_store(b)
    assert ha.source == "plain-table" and hb.source == "shadow-tables"
    assert np.array_equal(ha.rows, rows) and np.array_equal(hb.rows, rows)
    assert np.array_equal(ha.rowids, np.arange(5, 2505)) and np.array_equal(hb.rowids, ha.rowids)
    assert ha.file_paths == hb.file_paths == synth.default_paths(2500)
    assert ha.binary_count == 2500 and ha.vec0_count == 2500 and ha.dropped == 0


def test_loader_applies_inner_joins(tmp_path):
    """vec0 rows without an image_embeddings / images partner are not part of the
    reference's result set (image_database.py:1569-1570)."""
    rows = synth.unit_rows(300, DIM, 22)
    db = str(tmp_path / "o.db")
    synth.write_reference_db(db, rows, drop_mapping_for=[3, 100], drop_image_for=[10])
    h = loader.read_store(db)
    keep = np.delete(np.arange(300), [3, 10, 100])
    assert h.dropped == 3 and h.vec0_count == 300
    assert np.array_equal(h.rowids, keep + 1)
    assert np.array_equal(h.rows, rows[keep])
    # and that is exactly what SQLite returns for the statement
    conn, _ = sql_harness.connect(db)
    got = sql_harness.run_statement(conn, rows[0], -1, with_rowid=True)
    assert sorted(r[1] for r in got) == (keep + 1).tolist()
    conn.close()


def test_loader_incremental(tmp_path):
    rows = synth.unit_rows(120, DIM, 23)
    db = str(tmp_path / "i.db")
    synth.write_reference_db(db, rows)
    h = loader.read_store(db, min_rowid=100)
    assert np.array_equal(h.rowids, np.arange(101, 121)) and np.array_equal(h.rows, rows[100:])


def test_loader_rejects_wrong_dim(tmp_path):
    rows = synth.unit_rows(10, 64, 24)
    db = str(tmp_path / "d.db")
    synth.write_reference_db(db, rows)
    with pytest.raises(ValueError):
        loader.read_store(db, expect_dim=DIM)
    assert loader.read_store(db).dim == 64


@pytest.mark.parametrize("folders", [
    ["/data/photos/b"], ["/DATA/Photos/A/"], ["/data/100%_done"], ["/data/100x_done", "/data/scans"],
    ["/data/ph"], ["/nope"], ["/data/photos/é"], ["/data/photos/É"], ["/data/back\\slash"],
])
def test_folder_mask_equals_sqlite_like(tmp_path, folders):
    """like_prefix_mask == the rows SQLite's LIKE ... ESCAPE admits for the reference's
    pattern construction (ASCII-only case folding, escaped % and _)."""
    paths = synth.default_paths(60)
    paths[0] = "/data/100%_done/x.jpg"
    paths[1] = "/data/100x_done/x.jpg"
    paths[2] = "/data/100%xdone/x.jpg"
    paths[3] = "/data/photos/é/x.jpg"
    paths[4] = "/data/photos/É/x.jpg"
    paths[5] = "/DATA/PHOTOS/B/upper.jpg"
    paths[6] = "/data/photos/bb/x.jpg"
    paths[7] = "/data/back\\slash/x.jpg"
    paths[8] = "/data/ph"
    db = str(tmp_path / "f.db")
    conn = sqlite3.connect(db)
    conn.execute("CREATE TABLE images (id INTEGER PRIMARY KEY, file_path TEXT)")
    conn.executemany("INSERT INTO images VALUES (?, ?)", list(enumerate(paths)))
    where, params = sql_harness.where_clause_and_params(folders)
    got = {r[0] for r in conn.execute(f"SELECT i.id FROM images i {where}", params)}
    conn.close()
    mask = database.like_prefix_mask(paths, folders)
    assert set(np.nonzero(mask)[0].tolist()) == got


def test_duplicate_filter_equals_reference_restatement(tmp_path):
    rng = np.random.default_rng(31)
    rows = synth.unit_rows(400, DIM, 30)
    # families of near-identical sign codes: flip 0..4 signs of a parent row
    for child, parent, flips in [(10, 3, 0), (11, 3, 1), (12, 3, 2), (13, 3, 3), (50, 40, 2), (51, 50, 2),
                                 (52, 51, 2), (90, 80, 4)]:
        v = rows[parent].copy()
        idx = rng.choice(DIM, size=flips, replace=False)
        v[idx] = -v[idx]
        rows[child] = v
    db = str(tmp_path / "dup.db")
    synth.write_reference_db(db, rows)
    paths = synth.default_paths(400)
    for trial in range(20):
        pick = rng.permutation(120)[:40]
        sims = np.sort(rng.random(40))[::-1]
        if trial % 3 == 0:
            sims[5:9] = sims[5]                       # equal similarities
        if trial % 4 == 0:
            sims = rng.random(40)                     # unsorted input: exercises the replace branch
        results = [(paths[p], float(s)) for p, s in zip(pick, sims)]
        if trial % 5 == 0:
            results.append(("/not/in/db.jpg", 0.5))
        codes = {paths[p]: (rows[p] >= 0).astype(np.uint8) for p in pick}
        mine = database.filter_duplicates(list(results), codes, 2)
        theirs = sql_harness.reference_filter_duplicates(db, list(results), 2)
        assert mine == theirs


def test_unit_rows_are_unit_and_seeded():
    a, b = synth.unit_rows(50, DIM, 1), synth.unit_rows(50, DIM, 1)
    assert np.array_equal(a, b)
    assert np.allclose(np.linalg.norm(a, axis=1), 1.0, atol=1e-6)
    h = synth.fp16_normalised(a)
    assert np.all(np.abs(np.linalg.norm(h, axis=1) - 1.0) < 1e-3)


def test_read_codes_order_and_validation(tmp_path):
    """The sign-code loader runs the fallback's own statement: rows in binary_embeddings rowid
    order, INNER JOINed to images; a blob of the wrong width is refused."""
    import sqlite3
    from clip_database_b200 import loader
    rows = synth.unit_rows(50, 1152, 9)
    db = str(tmp_path / "codes.db")
    synth.write_reference_db(db, rows, vectors=False)
    host = loader.read_codes(db, expect_dim=1152)
    assert host.codes.shape == (50, 1152) and host.image_ids.tolist() == list(range(1, 51))
    assert np.array_equal(host.codes, (rows >= 0).astype(np.uint8))
    conn = sqlite3.connect(db)
    conn.execute("DELETE FROM images WHERE id = 7")              # orphaned code: dropped by the join
    conn.commit()
    assert loader.read_codes(db).image_ids.tolist() == [i for i in range(1, 51) if i != 7]
    conn.execute("UPDATE binary_embeddings SET embedding = ? WHERE image_id = 9", (b"\x01" * 100,))
    conn.commit()
    conn.close()
    with pytest.raises(ValueError):
        loader.read_codes(db, expect_dim=1152)


@pytest.mark.parametrize("layout", ["standin", "shadow"])
def test_sharded_loading_covers_the_store_exactly_once(tmp_path, layout):
    """Each rank of a row-sharded deployment reads only its rowid range; the ranges are contiguous,
    equal (+-1) in JOINED rows and their concatenation is the whole store."""
    from clip_database_b200 import loader
    rows = synth.unit_rows(103, 1152, 12)
    db = str(tmp_path / "s.db")
    synth.write_reference_db(db, rows, vec0_layout=layout, rowid_start=5, drop_mapping_for=[0, 50], drop_image_for=[102])
    whole = loader.read_store(db)
    assert whole.rows.shape[0] == 100
    for world in (1, 3, 8):
        parts = []
        for rank in range(world):
            lo, hi, n = loader.shard_rowid_range(db, rank, world)
            assert n == 100
            part = loader.read_store(db, min_rowid=lo, max_rowid=hi)
            assert abs(part.rows.shape[0] - 100 / world) < 1
            parts.append(part)
        assert np.array_equal(np.concatenate([p.rowids for p in parts]), whole.rowids)
        assert np.array_equal(np.concatenate([p.rows for p in parts]), whole.rows)
        assert sum((p.file_paths for p in parts), []) == whole.file_paths


def test_streamed_reading_is_chunked_and_consistent(tmp_path):
    """iter_store yields the same rows as read_store, chunk by chunk, skipping orphans inside and across chunks;
    plan_shards tiles the rowid axis from one snapshot."""
    n = 700
    rows = synth.unit_rows(n, 32, 5)
    db = str(tmp_path / "s.db")
    drop_m, drop_i = [0, 99, 100, 101, 699], [350]
    synth.write_reference_db(db, rows, drop_mapping_for=drop_m, drop_image_for=drop_i, rowid_start=10)
    whole = loader.read_store(db)
    keep = [i for i in range(n) if i not in drop_m + drop_i]
    assert whole.rowids.tolist() == [10 + i for i in keep] and np.array_equal(whole.rows, rows[keep])
    assert whole.vec0_count == n and whole.dropped == len(drop_m) + len(drop_i)
    assert whole.mtimes.shape == (len(keep),) and whole.mtimes[0] == 1.7e9 + keep[0]
    with loader.snapshot(db) as conn:
        st = loader.StoreStats()
        chunks = list(loader.iter_store(conn, chunk_rows=64, stats=st))
        assert all(len(c.rowids) <= 64 for c in chunks) and len(chunks) >= n // 64
        assert np.array_equal(np.concatenate([c.rowids for c in chunks]), whole.rowids)
        assert np.array_equal(np.concatenate([c.rows for c in chunks]), whole.rows)
        assert [p for c in chunks for p in c.file_paths] == whole.file_paths
        assert st.vec0_rows == n and st.joined_rows == len(keep) and st.dim == 32
        ranges, mapped = loader.plan_shards(conn, 3)
        assert mapped == len(keep) and ranges[0][0] is None and ranges[-1][1] is None
        got = [loader.stream_store(db, lambda c: None, min_rowid=lo, max_rowid=hi, conn=conn) for lo, hi in ranges]
        assert np.array_equal(np.concatenate([g.rowids for g in got]), whole.rowids)
        sizes = [len(g.rowids) for g in got]
        assert max(sizes) - min(sizes) <= 1
        some = [whole.rowids[3], whole.rowids[400], whole.rowids[401], 10 + 99]      # the last one is an orphan
        picked = list(loader.read_rows_by_rowid(conn, some))
        assert np.concatenate([c.rowids for c in picked]).tolist() == sorted(some[:3])
        mp = loader.read_mapping(conn)
        assert np.array_equal(mp.rowids, whole.rowids) and np.array_equal(mp.image_ids, whole.image_ids)


def test_data_version_moves_only_when_someone_else_commits(tmp_path):
    import sqlite3
    db = str(tmp_path / "v.db")
    synth.write_reference_db(db, synth.unit_rows(10, 16, 1))
    watch = loader.connect(db)
    v0 = loader.data_version(watch)
    assert loader.data_version(watch) == v0
    w = sqlite3.connect(db)
    w.execute("UPDATE images SET last_modified = last_modified + 1 WHERE id = 3")
    w.commit()
    w.close()
    assert loader.data_version(watch) != v0
    watch.close()


def test_native_readers_statement_needs_no_sorter(tmp_path):
    """The native loader's statement (csrc/clipdb.cu: SQL_LOAD_ROWS) must walk vec0 in rowid order with two primary-key
    probes per row — the reference statement's own plan — and no temp b-tree (which would hold every blob)."""
    import os
    import re
    import sqlite3
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "clip_database_b200", "csrc",
                            "clipdb.cu")).read()
    body = re.search(r"SQL_LOAD_ROWS =\s*((?:\s*\"[^\"]*\"\s*)+);", src).group(1)
    sql = "".join(re.findall(r"\"([^\"]*)\"", body))
    assert "CROSS JOIN image_embeddings" in sql and "ORDER BY v.rowid" in sql
    db = str(tmp_path / "p.db")
    synth.write_reference_db(db, synth.unit_rows(50, 16, 2), drop_mapping_for=[3])
    conn = sqlite3.connect(db)
    plan = [r[-1] for r in conn.execute("EXPLAIN QUERY PLAN " + sql.replace("?1", "?").replace("?2", "?"), (0, 1 << 40))]
    assert plan[0].startswith("SEARCH v USING INTEGER PRIMARY KEY") and len(plan) == 3, plan
    assert not any("TEMP B-TREE" in p for p in plan), plan
    got = conn.execute(sql.replace("?1", "?").replace("?2", "?"), (0, 1 << 40)).fetchall()
    want = loader.read_store(db)
    assert [r[0] for r in got] == want.rowids.tolist() and [r[3] for r in got] == want.file_paths
    conn.close()
