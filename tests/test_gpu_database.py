"""``ImageDatabase`` as a long-lived process next to a scanner that keeps writing: ``refresh()`` must leave the
resident store answering exactly what the reference answers by reopening SQLite for every search
(image_database.py:1475) — checked against the reference's own SQL statement executed by the real SQLite on
the same file (oracle/sql_harness.py).  Also the streaming loader at 1M rows and the interactive session."""
import os
import resource
import sqlite3
import subprocess
import sys

import numpy as np
import pytest

from clip_database_b200 import synth
from oracle import sql_harness

from conftest import ROOT, have_gpu, tol

pytestmark = pytest.mark.gpu
DIM = 1152


def reference_answer(db_path, q, k, folders=None):
    return sql_harness.reference_search(db_path, q, k, folders)


def assert_same_answer(got, want):
    assert [p for p, _ in got] == [p for p, _ in want]
    g, w = np.array([s for _, s in got]), np.array([s for _, s in want])
    assert np.all(np.abs(g - w) <= tol(1.0 - w))


def rescan_modified_file(conn, file_path, new_embedding, mtime):
    """What ``_commit_batch`` does for a file whose content changed (image_database.py:1137-1198): the image row
    is re-keyed by INSERT OR REPLACE, a NEW vec0 row is inserted and linked; the old vec0 row stays, orphaned."""
    cur = conn.cursor()
    cur.execute("INSERT OR REPLACE INTO images (file_path, last_modified, file_hash) VALUES (?, ?, ?)",
                (file_path, mtime, "changed"))
    image_id = cur.lastrowid
    assert cur.execute("SELECT rowid FROM image_embeddings WHERE image_id = ?", (image_id,)).fetchone() is None
    cur.execute("INSERT INTO vec0 (embedding) VALUES (?)", (np.asarray(new_embedding, dtype=np.float32).tobytes(),))
    vec_rowid = cur.lastrowid
    cur.execute("INSERT INTO image_embeddings (rowid, image_id) VALUES (?, ?)", (vec_rowid, image_id))
    cur.execute("INSERT INTO binary_embeddings (image_id, embedding) VALUES (?, ?)",
                (image_id, (np.asarray(new_embedding) >= 0).astype(np.uint8).tobytes()))
    conn.commit()
    return image_id, vec_rowid


@pytest.mark.parametrize("devices,batch_store,budget", [(None, False, None), (None, True, None), ([0, 0], False, None),
                                                        (None, True, 12_000_000)],
                         ids=["one-gpu", "one-gpu-bf16", "two-shards", "one-gpu-tiered"])
def test_refresh_tracks_a_scanner_writing_next_to_it(tmp_path, devices, batch_store, budget):
    assert have_gpu(), "GPU tests selected but no CUDA device is visible"
    from clip_database_b200 import ImageDatabase
    n, k = 3000, 12
    rows = synth.unit_rows(n, DIM, 31)
    paths = synth.default_paths(n)
    db_path = str(tmp_path / "live.db")
    synth.write_reference_db(db_path, rows, paths, drop_mapping_for=[17])
    q = synth.unit_rows(1, DIM, 32)[0]
    db = ImageDatabase(db_path, device=0, devices=devices, batch_store=batch_store, hbm_budget_bytes=budget)
    # a 12 MB "GPU": 13.8 MB of float32 + 6.9 MB of bf16 do not fit -> bf16 copy + 768 float32 rows resident, the
    # other float32 rows in pinned host memory, every search through the pre-selection + exact re-rank
    assert db.placement == ("tiered" if budget else "device")
    writer = sqlite3.connect(db_path)
    try:
        first = db.search_embedding(q, k=k, show_duplicates=True)
        assert_same_answer(first, reference_answer(db_path, q, k))
        assert db.refresh() == 0                                      # nothing was committed: O(1)

        # 1. a modified file is re-scanned: its old row must disappear, its new embedding must be found
        top_path = first[0][0]
        new_vec = synth.unit_rows(1, DIM, 33)[0]
        rescan_modified_file(writer, top_path, new_vec, 1.8e9)
        assert db.refresh() == 2                                      # one row retired, one appended
        after = db.search_embedding(q, k=k, show_duplicates=True)
        assert_same_answer(after, reference_answer(db_path, q, k))
        assert top_path not in [p for p, _ in after]                  # (its new embedding is unrelated to q)
        near_new = db.search_embedding(new_vec, k=3, show_duplicates=True)
        assert near_new[0][0] == top_path and abs(near_new[0][1] - 1.0) < 1e-5
        assert_same_answer(near_new, reference_answer(db_path, new_vec, 3))

        # 2. plain appends (new files)
        more = synth.unit_rows(40, DIM, 34)
        more[5] = q                                                   # one of them is the query itself
        cur = writer.cursor()
        for j in range(40):
            cur.execute("INSERT INTO images (file_path, last_modified, file_hash) VALUES (?, ?, ?)",
                        (f"/data/new/img_{j:04d}.jpg", 1.9e9 + j, "x"))
            image_id = cur.lastrowid
            cur.execute("INSERT INTO vec0 (embedding) VALUES (?)", (more[j].tobytes(),))
            cur.execute("INSERT INTO image_embeddings (rowid, image_id) VALUES (?, ?)", (cur.lastrowid, image_id))
        writer.commit()
        assert db.refresh() == 40
        got = db.search_embedding(q, k=k, show_duplicates=True)
        assert got[0][0] == "/data/new/img_0005.jpg"
        assert_same_answer(got, reference_answer(db_path, q, k))

        # 3. an in-place re-embedding (UPDATE vec0 ... WHERE rowid, :1165-1167) with the image row touched
        victim_path = got[1][0]
        image_id, vec_rowid = writer.execute(
            "SELECT i.id, ie.rowid FROM images i JOIN image_embeddings ie ON ie.image_id = i.id WHERE i.file_path = ?",
            (victim_path,)).fetchone()
        writer.execute("UPDATE vec0 SET embedding = ? WHERE rowid = ?", (synth.unit_rows(1, DIM, 35)[0].tobytes(), vec_rowid))
        writer.execute("UPDATE images SET last_modified = ? WHERE id = ?", (2.0e9, image_id))
        writer.commit()
        assert db.refresh() == 1
        got = db.search_embedding(q, k=k, show_duplicates=True)
        assert victim_path not in [p for p, _ in got]
        assert_same_answer(got, reference_answer(db_path, q, k))

        # 4. a folder filter on top of retired rows
        folders = ["/data/photos/a", "/data/new"]
        assert_same_answer(db.search_embedding(q, k=k, filter_folders=folders, show_duplicates=True),
                           reference_answer(db_path, q, k, folders))

        # 5. a mapping appears for a vec0 row that had none at load time: a full reload
        writer.execute("INSERT INTO image_embeddings (rowid, image_id) VALUES (?, ?)", (18, 18))
        writer.commit()
        assert db.refresh() > 40
        assert_same_answer(db.search_embedding(rows[17], k=3, show_duplicates=True), reference_answer(db_path, rows[17], 3))
        assert db.refresh() == 0
    finally:
        writer.close()
        db.close()


def test_session_answers_vector_queries_and_uses_the_embedder_hook(tmp_path):
    assert have_gpu()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from clip_database_b200 import ImageDatabase, session
    from fake_embedder import HashEmbedder
    n = 2000
    rows = synth.unit_rows(n, DIM, 41)
    db_path = str(tmp_path / "s.db")
    synth.write_reference_db(db_path, rows)
    emb = HashEmbedder()
    vq, vs, vn = synth.unit_rows(3, DIM, 42)
    for name, v in (("q", vq), ("s", vs), ("n", vn)):
        np.save(tmp_path / f"{name}.npy", v)
    img = str(tmp_path / "y.jpg")
    open(img, "w").close()
    with _closing(ImageDatabase(db_path, device=0)) as db:
        want_vec = db.search_embedding(vq, k=5, embedding2=vs, negative_embeddings=[vn], negative_weights=[0.5])
        want_txt = db.search_embedding(emb.text("a red car"), k=5, embedding2=emb.image(img))
        want_mix = db.search_embedding(emb.text("boats"), k=5, negative_embeddings=[vn], negative_weights=[0.5])

    def scripted(lines):
        it = iter(lines)
        out = []

        def read(_prompt):
            try:
                return next(it)
            except StopIteration:
                raise EOFError
        return read, out.append, out

    def listed(out):
        return [ln.split(": ", 1)[1] for ln in out if ln.startswith("  ") and ". " in ln and ": " in ln and "/" in ln]

    # no embedder: vector queries only
    read, write, out = scripted(["k:5", f"vector:{tmp_path}/q.npy + vector:{tmp_path}/s.npy - vector:{tmp_path}/n.npy",
                                 "a red car", "quit"])
    assert session.main(["--db", db_path], read=read, write=write) == 0
    assert listed(out) == [p for p, _ in want_vec]
    assert any("no embedder configured" in ln for ln in out)          # the text query is refused, the session goes on
    # with the embedder hook: text, image: and vector: in one session
    read, write, out = scripted(["k:5", f"a red car + image:{img}", "quit"])
    assert session.main(["--db", db_path, "--embedder", "fake_embedder:HashEmbedder"], read=read, write=write) == 0
    assert listed(out) == [p for p, _ in want_txt]
    read, write, out = scripted(["k:5", f"boats - vector:{tmp_path}/n.npy", "quit"])
    assert session.main(["--db", db_path, "--embedder", "fake_embedder:HashEmbedder"], read=read, write=write) == 0
    assert listed(out) == [p for p, _ in want_mix]


class _closing:
    def __init__(self, db):
        self.db = db

    def __enter__(self):
        return self.db

    def __exit__(self, *exc):
        self.db.close()


MAKE_BIG_DB = r"""
import sys
sys.path.insert(0, {root!r})
from clip_database_b200 import synth
from oracle import ref
rows = ref.fill_unit_rows({n}, 1152, 1234)
synth.write_reference_db({path!r}, rows, binary_codes=False)
import sqlite3
c = sqlite3.connect({path!r})
c.execute("INSERT INTO binary_embeddings (image_id, embedding) VALUES (1, zeroblob(1152))")   # the guard at :1488-1500
c.commit()
"""


def test_loader_streams_a_million_rows(tmp_path):
    """SURVEY §8d config 1 says "loaded to HBM through the loader"; VERDICT r1 weak #6: the loader must survive a
    real database.  1M rows (4.6 GB of blobs) written by a child process, streamed into HBM here: host memory
    grows by far less than the store, rows/s is reported, and the answers equal the reference statement's."""
    assert have_gpu()
    import shutil
    import time
    from clip_database_b200 import ImageDatabase
    n = 1_000_000
    if shutil.disk_usage(str(tmp_path)).free < 12e9:
        pytest.skip("needs ~6 GB of scratch disk")
    db_path = str(tmp_path / "big.db")
    subprocess.run([sys.executable, "-c", MAKE_BIG_DB.format(root=ROOT, n=n, path=db_path)], check=True, timeout=900)
    rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss          # KiB
    t0 = time.perf_counter()
    db = ImageDatabase(db_path, device=0, verbose=True)
    secs = time.perf_counter() - t0
    rss1 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    try:
        assert db.index.num_rows == n
        grown = (rss1 - rss0) * 1024
        print(f"loader: {n / secs:.0f} rows/s ({n * 4608 / 1e6 / secs:.0f} MB/s), host RSS grew by {grown / 1e6:.0f} MB "
              f"for a {n * 4608 / 1e6:.0f} MB store")
        assert grown < 1.5e9, "the loader must not hold the store in host memory"
        for seed in (7, 8):
            q = synth.unit_rows(1, DIM, seed)[0]
            assert_same_answer(db.search_embedding(q, k=20, show_duplicates=True), reference_answer(db_path, q, 20))
    finally:
        db.close()


def test_a_store_too_large_for_the_gpu_says_what_to_do(tmp_path):
    assert have_gpu()
    from clip_database_b200 import ImageDatabase
    db_path = str(tmp_path / "big.db")
    synth.write_reference_db(db_path, synth.unit_rows(3000, DIM, 61))
    with pytest.raises(MemoryError, match="batch_store=True"):
        ImageDatabase(db_path, device=0, hbm_budget_bytes=12_000_000)
    with pytest.raises(MemoryError, match="devices="):
        ImageDatabase(db_path, device=0, batch_store=True, hbm_budget_bytes=5_000_000)


def test_a_mapping_without_a_vector_does_not_cause_reload_after_reload(tmp_path):
    """A dangling image_embeddings row (its vec0 row is missing) can never be loaded; refresh() must notice once
    and then leave the store alone."""
    assert have_gpu()
    from clip_database_b200 import ImageDatabase
    rows = synth.unit_rows(500, DIM, 71)
    db_path = str(tmp_path / "d.db")
    synth.write_reference_db(db_path, rows)
    w = sqlite3.connect(db_path)
    w.execute("DELETE FROM vec0 WHERE rowid = 100")
    w.commit()
    with _closing(ImageDatabase(db_path, device=0)) as db:
        assert db.index.num_rows == 499 and db.reloads == 1
        for step in range(3):
            w.execute("UPDATE images SET file_hash = ? WHERE id = 7", (f"h{step}",))     # unrelated commits
            w.commit()
            db.refresh()
            assert db.reloads == 2, "one reload to learn that rowid 100 is permanently unloadable, then none"
        q = rows[99]                                                                      # the row that is missing
        assert_same_answer(db.search_embedding(q, k=5, show_duplicates=True), reference_answer(db_path, q, 5))
    w.close()


def test_native_and_python_readers_load_the_same_store(tmp_path):
    """clipdb_append_sqlite (SQLite's C library -> pinned double buffer -> DMA) against the Python reader: same rows
    in the same order, same metadata (orphans dropped, unicode paths, ranges), same answers; sqlite-vec style
    shadow tables fall back to the Python reader."""
    assert have_gpu()
    from clip_database_b200 import GpuIndex, ImageDatabase, loader
    n = 20_000                                                     # more than two 8192-row chunks
    rows = synth.unit_rows(n, DIM, 81)
    paths = synth.default_paths(n)
    paths[3] = "/data/photos/a/naïve – café ☕.jpg"
    paths[19_999] = "/data/scans/日本語 ファイル.png"
    db_path = str(tmp_path / "n.db")
    synth.write_reference_db(db_path, rows, paths, rowid_start=5, drop_mapping_for=[0, 8191, 8192, 12_345],
                             drop_image_for=[77, 19_998])
    q = synth.unit_rows(3, DIM, 82)
    with _closing(ImageDatabase(db_path, device=0, native_loader=False)) as py, \
            _closing(ImageDatabase(db_path, device=0)) as nat:
        assert py.load_source == "plain-table" and nat.load_source == "plain-table (native reader)"
        assert nat.index.num_rows == py.index.num_rows == n - 6
        assert np.array_equal(nat._rowids, py._rowids) and np.array_equal(nat._image_ids, py._image_ids)
        assert np.array_equal(nat._mtimes, py._mtimes) and nat._paths == py._paths
        assert nat._vec0_count == py._vec0_count == n
        for v in q:
            assert nat.search_embedding(v, k=15, show_duplicates=True) == py.search_embedding(v, k=15, show_duplicates=True)
            assert_same_answer(nat.search_embedding(v, k=15, show_duplicates=True), reference_answer(db_path, v, 15))
        hit = nat.search_embedding(rows[3], k=1, show_duplicates=True)
        assert hit[0][0] == paths[3]
    # a rowid range through the raw entry point, with the metadata callback
    with GpuIndex(0) as idx:
        idx.reserve(100, DIM, explicit_rowids=True)
        seen = []
        vec0_rows, joined = idx.append_sqlite(db_path, 5 + 8000, 5 + 8400, lambda *a: seen.append(a), chunk_rows=128)
        assert vec0_rows == 400 and joined == 398 and idx.num_rows == 398          # positions 8191, 8192 are orphans
        assert sum(len(a[0]) for a in seen) == 398 and len(seen) == 4
        assert seen[0][0][0] == 5 + 8001 and seen[0][3][0] == paths[8001]
        res = idx.search(rows[8100], 1)
        assert res.rowids[0, 0] == 5 + 8100
        with pytest.raises(Exception, match="float32"):                            # a store of another dimension
            with GpuIndex(0) as other:
                other.reserve(10, 64, explicit_rowids=True)
                other.append_sqlite(db_path)
    # shadow tables: not a plain table -> the Python reader
    shadow = str(tmp_path / "s.db")
    synth.write_reference_db(shadow, rows[:3000], vec0_layout="shadow")
    with _closing(ImageDatabase(shadow, device=0)) as db:
        assert db.load_source == "shadow-tables" and db.index.num_rows == 3000
        assert db.search_embedding(rows[10], k=1, show_duplicates=True)[0][0] == synth.default_paths(3000)[10]


def test_tiered_shards_on_several_gpus(tmp_path):
    """ImageDatabase(devices=[0, 1], batch_store=True) with shards that do not fit their GPU: both shards are loaded
    tiered and every search goes through the bf16 pre-selection + exact re-rank with the candidates exchanged by the
    batched path's last kernel.  Needs one GPU per shard."""
    assert have_gpu()
    import torch
    from clip_database_b200 import ImageDatabase
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    n, k = 6000, 10
    rows = synth.unit_rows(n, DIM, 91)
    db_path = str(tmp_path / "t.db")
    synth.write_reference_db(db_path, rows, drop_mapping_for=[5, 4000])
    # per shard: 3000 rows = 13.8 MB float32 + 6.9 MB bf16 against a 12 MB budget -> tiered
    with _closing(ImageDatabase(db_path, devices=[0, 1], batch_store=True, hbm_budget_bytes=12_000_000)) as db:
        assert db.placement == "tiered" and db.index.prefer_batch
        for seed in (92, 93, 94):
            q = synth.unit_rows(1, DIM, seed)[0]
            before = db.index.launch_count
            got = db.search_embedding(q, k=k, show_duplicates=True)
            assert db.index.launch_count - before <= 2 * 8, "the batched path was not taken"
            assert_same_answer(got, reference_answer(db_path, q, k))
        folders = ["/data/photos/a"]                              # a mask: the exact (fused) path over the tiers
        assert_same_answer(db.search_embedding(rows[9], k=k, filter_folders=folders, show_duplicates=True),
                           reference_answer(db_path, rows[9], k, folders))
        many = db.search_embeddings(rows[100:140], k=5)
        assert [r[0][0] for r in many] == [synth.default_paths(n)[i] for i in range(100, 140)]


@pytest.mark.parametrize("readers", [1, 4])
def test_native_reader_on_a_sparse_rowid_space_and_with_many_readers(tmp_path, readers):
    """Stripes are ranges of rowid VALUES; a table whose rowids are mostly gaps is read as one ordered pass instead,
    and the answer does not depend on how many connections read the stripes."""
    assert have_gpu()
    import sqlite3 as sq3
    from clip_database_b200 import GpuIndex, loader
    n = 5000
    rows = synth.unit_rows(n, DIM, 95)
    dense = str(tmp_path / "dense.db")
    synth.write_reference_db(dense, rows, rowid_start=1000, drop_mapping_for=[0, 127, 128, 4999])
    sparse = str(tmp_path / "sparse.db")
    synth.write_reference_db(sparse, rows, rowid_start=1000)
    w = sq3.connect(sparse)
    w.execute("DELETE FROM vec0 WHERE (rowid - 1000) % 10 != 0")           # 500 rows over a span of 5000 rowids
    w.commit()
    w.close()
    for path in (dense, sparse):
        want = loader.read_store(path)
        with GpuIndex(0) as idx:
            idx.set_option("sqlite_readers", readers)
            idx.reserve(len(want.rowids) + 10, DIM, explicit_rowids=True)
            got_ids, got_paths, chunks = [], [], []
            vec0_rows, joined = idx.append_sqlite(path, on_chunk=lambda i, im, mt, p: (got_ids.append(i), got_paths.extend(p),
                                                                                      chunks.append(len(i))),
                                                  chunk_rows=128)
            assert joined == len(want.rowids) == idx.num_rows and vec0_rows == want.vec0_count
            assert np.array_equal(np.concatenate(got_ids), want.rowids) and got_paths == want.file_paths
            assert max(chunks) <= 128
            with GpuIndex(0) as ref_idx:
                ref_idx.load(want.rows, want.rowids)
                q = synth.unit_rows(4, DIM, 96)
                a, b = idx.search(q, 10), ref_idx.search(q, 10)
                assert np.array_equal(a.rowids, b.rowids) and np.array_equal(a.distances.view(np.uint32), b.distances.view(np.uint32))
