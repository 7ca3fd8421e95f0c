#!/usr/bin/env python3
"""Generate tests/golden/reference_rescan.json: what the REFERENCE's own writer does to the database when files are
(re)scanned, and what the reference's own search() returns afterwards.

Run in the authoring container only (needs /root/reference).  What executes: ``image_database.py`` imported
unmodified (stub ``sqlite_vec`` as in make_golden.py); the database starts as a synthetic reference-schema file
(``synth.write_reference_db``, plain stand-in ``vec0``); each stage hands a batch of
``(file_path, last_modified, file_hash, embedding)`` to the reference's ``ImageDatabase._commit_batch``
(image_database.py:1098-1204) with ``save_full_embeddings=True`` and commits, then runs the reference's ``search()``.

Recorded per stage: the batch (embedding seeds), the row-level difference the writer made to ``images``,
``image_embeddings`` and ``vec0`` (so a test can replay exactly those writes with plain SQL where the reference is
absent), and the search results.  This pins how a resident copy must follow the file: a modified file is re-keyed by
``INSERT OR REPLACE INTO images`` (its old vec0 row stays, orphaned, and the INNER JOINs drop it), an unchanged
file is skipped, a new file is appended.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sqlite3
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from clip_database_b200 import synth  # noqa: E402
import make_golden  # noqa: E402

OUT = os.path.join(HERE, "reference_rescan.json")
DIM = 1152
N = 1500
ROWS_SEED, QUERY_SEED = 301, 302


def snapshot(conn):
    images = {r[0]: (r[1], r[2], r[3]) for r in conn.execute("SELECT id, file_path, last_modified, file_hash FROM images")}
    mapping = {r[0]: r[1] for r in conn.execute("SELECT rowid, image_id FROM image_embeddings")}
    vec0 = {r[0]: r[1] for r in conn.execute("SELECT rowid, embedding FROM vec0")}
    return images, mapping, vec0


def diff(before, after):
    out = {}
    for name, b, a in zip(("images", "image_embeddings", "vec0"), before, after):
        out[name] = {"deleted": sorted(k for k in b if k not in a),
                     "upserted": sorted(k for k in a if k not in b or b[k] != a[k])}
    return out


def stages(paths, top_paths):
    """[(name, [(file_path, last_modified, file_hash, embedding seed)])]"""
    return [
        ("modified_file_and_new_files",
         [(top_paths[0], 1.8e9, "changed-0", 401),            # the best hit's file was edited: re-keyed, re-embedded
          ("/data/new/first.jpg", 1.8e9 + 1, "new-0", 402),
          ("/data/new/second.jpg", 1.8e9 + 2, "new-1", QUERY_SEED)]),   # a new file whose embedding IS the query
        ("unchanged_file_is_skipped_another_is_modified",
         [(top_paths[0], 1.8e9, "changed-0", 403),            # same path, same mtime: already processed -> skipped
          (top_paths[1], 1.9e9, "changed-1", 404)]),
        ("same_file_modified_twice",
         [(top_paths[1], 2.0e9, "changed-2", 405),
          ("/data/new/first.jpg", 2.0e9, "new-0b", 406)]),
    ]


def main():
    idb = make_golden.import_reference()
    rows = synth.unit_rows(N, DIM, ROWS_SEED)
    q = synth.unit_rows(1, DIM, QUERY_SEED)[0]
    paths = synth.default_paths(N)
    tmp = tempfile.mkdtemp(prefix="golden_rescan_")
    db_path = os.path.join(tmp, "rescan.db")
    synth.write_reference_db(db_path, rows, paths)
    db = object.__new__(idb.ImageDatabase)
    db.db_path = db_path
    db.embedding_dim = DIM
    db._get_text_embedding = lambda text: q
    db._get_image_embedding = lambda path: q

    def search():
        with contextlib.redirect_stdout(io.StringIO()):
            return db.search("q1", k=10, show_duplicates=True)

    first = search()
    out = {"generated_by": "reference image_database.py ImageDatabase._commit_batch() + search(), sqlite_vec stubbed",
           "n": N, "rows_seed": ROWS_SEED, "query_seed": QUERY_SEED, "dim": DIM, "k": 10,
           "initial": {"paths": [p for p, _ in first], "similarities": [float(s) for _, s in first]}, "stages": []}
    top_paths = [p for p, _ in first]
    conn = sqlite3.connect(db_path)
    for name, batch in stages(paths, top_paths):
        before = snapshot(conn)
        cur = conn.cursor()
        with contextlib.redirect_stdout(io.StringIO()):
            db._commit_batch(cur, [(fp, mt, fh, synth.unit_rows(1, DIM, seed)[0]) for fp, mt, fh, seed in batch], True)
        conn.commit()
        after = snapshot(conn)
        d = diff(before, after)
        seed_of = {synth.unit_rows(1, DIM, seed)[0].tobytes(): seed for _, _, _, seed in batch}
        rec = {"name": name, "batch": [[fp, mt, fh, seed] for fp, mt, fh, seed in batch],
               "images": {"deleted": d["images"]["deleted"],
                          "upserted": [[k, *after[0][k]] for k in d["images"]["upserted"]]},
               "image_embeddings": {"deleted": d["image_embeddings"]["deleted"],
                                    "upserted": [[k, after[1][k]] for k in d["image_embeddings"]["upserted"]]},
               "vec0": {"deleted": d["vec0"]["deleted"],
                        "upserted": [[k, seed_of[after[2][k]]] for k in d["vec0"]["upserted"]]}}
        res = search()
        rec["paths"] = [p for p, _ in res]
        rec["similarities"] = [float(s) for _, s in res]
        out["stages"].append(rec)
        print(name, {k: (len(v["deleted"]), len(v["upserted"])) for k, v in rec.items() if isinstance(v, dict)},
              "->", res[:2])
    conn.close()
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
