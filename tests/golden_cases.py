"""Rebuild the inputs of a golden case (shared by the CPU and GPU golden tests)."""
import hashlib
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(_HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


def inputs_for(case):
    """(rows, paths, search kwargs, named vectors, drop_mapping, drop_image), verified
    against the SHA-256 recorded when the reference produced the expected output."""
    rows, paths, kwargs, vectors = make_golden.materialise_case(case)
    rows, drop_m, drop_i = make_golden.post_adjust(case, rows, vectors["q1"])
    digest = hashlib.sha256(np.ascontiguousarray(rows).tobytes()).hexdigest()
    assert digest == case["rows_sha256"], "seeded inputs differ from the ones the golden file was made with"
    assert drop_m == case["drop_mapping_for"] and drop_i == case["drop_image_for"]
    return rows, paths, kwargs, vectors, drop_m, drop_i


def embedding_call(kwargs, vectors):
    """Translate the reference-style search kwargs (names of vectors) into the
    arguments of ``search_embedding`` / the oracle's ``compose_query``."""
    e1 = vectors["q1"]
    e2 = vectors[kwargs["query2"]] if "query2" in kwargs else None
    weights = kwargs.get("weights", (0.5, 0.5))
    negs, ws = [], []
    if "negative_query" in kwargs:
        negs.append(vectors[kwargs["negative_query"]])
        ws.append(kwargs.get("negative_weight", 0.5))
    for i, name in enumerate(kwargs.get("negative_queries", [])):
        negs.append(vectors[name])
        nw = kwargs.get("negative_weights")
        ws.append(nw[i] if nw and i < len(nw) else kwargs.get("negative_weight", 0.5))
    return e1, e2, weights, negs, ws


def rescan_golden():
    import json
    with open(os.path.join(_HERE, "golden", "reference_rescan.json")) as f:
        return json.load(f)


def apply_rescan_stage(conn, stage, dim=1152):
    """Replay, with plain SQL, exactly the row-level writes the reference's own ``_commit_batch`` made in this stage
    (recorded by tests/golden/make_golden_rescan.py): deleted / upserted rows of images, image_embeddings, vec0."""
    from clip_database_b200 import synth
    cur = conn.cursor()
    for image_id in stage["images"]["deleted"]:
        cur.execute("DELETE FROM images WHERE id = ?", (image_id,))
    for image_id, file_path, mtime, file_hash in stage["images"]["upserted"]:
        cur.execute("INSERT OR REPLACE INTO images (id, file_path, last_modified, file_hash) VALUES (?, ?, ?, ?)",
                    (image_id, file_path, mtime, file_hash))
    for rowid in stage["vec0"]["deleted"]:
        cur.execute("DELETE FROM vec0 WHERE rowid = ?", (rowid,))
    for rowid, seed in stage["vec0"]["upserted"]:
        cur.execute("INSERT OR REPLACE INTO vec0 (rowid, embedding) VALUES (?, ?)",
                    (rowid, synth.unit_rows(1, dim, seed)[0].tobytes()))
    for rowid in stage["image_embeddings"]["deleted"]:
        cur.execute("DELETE FROM image_embeddings WHERE rowid = ?", (rowid,))
    for rowid, image_id in stage["image_embeddings"]["upserted"]:
        cur.execute("INSERT OR REPLACE INTO image_embeddings (rowid, image_id) VALUES (?, ?)", (rowid, image_id))
    conn.commit()
