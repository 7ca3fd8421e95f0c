#!/usr/bin/env python3
"""Generate tests/golden/reference_binary_search.json by running the REFERENCE's own
``ImageDatabase.search()`` on BINARY-ONLY databases (``vec0`` empty), which makes it take its
sign-code fallback (image_database.py:1591-1629).

Same harness as make_golden.py (reference imported unmodified, ``sqlite_vec`` stubbed, the
embedding methods replaced by table look-ups); nothing of sqlite-vec's arithmetic is involved
on this path, so these outputs are pinned by the reference + numpy + the real SQLite alone —
including numpy's uint8 wrap-around of the score and SQLite's choice of scan order for the
filtered statement.  Run in the authoring container only (needs /root/reference).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden  # noqa: E402
from clip_database_b200 import synth  # noqa: E402

OUT = os.path.join(HERE, "reference_binary_search.json")
DIM = 1152


def scrambled_paths(n):
    """file_path order differs from id order inside every folder, so the scan order SQLite
    picks for the filtered statement (file_path index vs binary_embeddings rowid) shows up
    in the tie order."""
    folders = ("/data/photos/a", "/data/photos/b", "/data/scans")
    return [f"{folders[i % 3]}/img_{(i * 7919) % 100003:08d}.jpg" for i in range(n)]


def build_cases():
    cases = []

    def add(name, **kw):
        kw["name"] = name
        cases.append(kw)

    base = dict(n=6000, rows_seed=1234, query_seed=99)
    add("binary_k20", **base, k=20, show_duplicates=True)
    add("binary_k1", **base, k=1, show_duplicates=True)
    add("binary_k200", **base, k=200, show_duplicates=True)
    add("binary_duplicate_filter", **base, k=20, show_duplicates=False)
    add("binary_folder_filter", **base, k=20, show_duplicates=True, filter_folders=["/data/photos/b"],
        scrambled=True)
    add("binary_two_folders", **base, k=50, show_duplicates=True, filter_folders=["/DATA/photos/a", "/data/scans/"],
        scrambled=True)
    add("binary_blend_07_03_negative", **base, k=20, show_duplicates=True, query2_seed=100,
        weights=[0.7, 0.3], negative_seeds=[101], negative_weights=[0.5], legacy_negative=True)
    add("binary_three_negatives", **base, k=20, show_duplicates=True, negative_seeds=[101, 102, 103],
        negative_weights=[0.5, 0.25, 1.5])
    add("binary_k_zero", n=300, rows_seed=7, query_seed=99, k=0, show_duplicates=True)
    add("binary_k_exceeds_rows", n=300, rows_seed=7, query_seed=99, k=350, show_duplicates=True)
    add("binary_k_negative_is_a_python_slice", n=300, rows_seed=7, query_seed=99, k=-280, show_duplicates=True)
    add("binary_k_negative_filtered", n=300, rows_seed=7, query_seed=99, k=-90, show_duplicates=True,
        filter_folders=["/data/photos/a"], scrambled=True)
    return cases


def paths_for(case):
    return scrambled_paths(case["n"]) if case.get("scrambled") else synth.default_paths(case["n"])


def main():
    idb = make_golden.import_reference()
    tmp = tempfile.mkdtemp(prefix="golden_bin_")
    out = {"generated_by": "reference image_database.py ImageDatabase.search() on binary-only databases "
                           "(vec0 empty -> sign-code fallback, image_database.py:1591-1629)",
           "numpy": np.__version__, "sqlite": __import__("sqlite3").sqlite_version, "dim": DIM, "cases": []}
    for case in build_cases():
        rows, _, kwargs, vectors = make_golden.materialise_case(case)
        paths = paths_for(case)
        db_path = os.path.join(tmp, case["name"] + ".db")
        synth.write_reference_db(db_path, rows, paths, vectors=False)
        db = object.__new__(idb.ImageDatabase)
        db.db_path = db_path
        db.embedding_dim = DIM
        db._get_text_embedding = lambda text, _v=vectors: _v[text]
        db._get_image_embedding = lambda path, _v=vectors: _v[path]
        with contextlib.redirect_stdout(io.StringIO()):
            results = db.search("q1", **kwargs)
        pos = {p: i for i, p in enumerate(paths)}
        rec = dict(case)
        rec["rows_sha256"] = make_golden.sha(rows)
        rec["expected_positions"] = [pos[p] for p, _ in results]
        rec["expected_similarities"] = [float(s) for _, s in results]
        out["cases"].append(rec)
        print(f"{case['name']:40s} -> {len(results)} results, first {results[:2]}")
        os.remove(db_path)
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
