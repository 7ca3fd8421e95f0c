"""Sign-code fallback search on the GPU through the C ABI: bit-exact
against the oracle (integer work) and against the outputs of the reference's own code."""
import json
import os

import numpy as np
import pytest

from clip_database_b200 import synth
from oracle import binary as obinary

from conftest import have_gpu
import test_binary_cpu as cpu

pytestmark = pytest.mark.gpu
DIM = 1152


@pytest.fixture(scope="module")
def gpu():
    if not have_gpu():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    from clip_database_b200 import GpuIndex
    return GpuIndex


@pytest.fixture(scope="module")
def store(gpu):
    rng = np.random.default_rng(1234)
    codes = (rng.standard_normal((100_003, DIM), dtype=np.float32) >= 0).astype(np.uint8)
    codes[17] = 1                      # all ones: popcount = |query|, wraps modulo 256
    codes[50_000] = codes[17]
    queries = (rng.standard_normal((8, DIM), dtype=np.float32) >= 0).astype(np.uint8)
    queries[7] = 1                     # score = row popcount (~576): far above 255
    idx = gpu(0)
    idx.load_codes(codes)
    yield codes, queries, idx
    idx.close()


@pytest.mark.parametrize("mode,wrap", [("reference", True), ("popcount", False)])
@pytest.mark.parametrize("k", [1, 20, 33, 100, 128, 129, 1000])
def test_matches_oracle(store, mode, wrap, k):
    codes, queries, idx = store
    for q in queries[[0, 3, 7]]:
        pos, score = idx.binary_search(q, k, score_mode=mode)
        opos, oscore = obinary.search(codes, q, k, wrap=wrap)
        assert np.array_equal(pos, opos)
        assert np.array_equal(score, oscore)


def test_wraparound_is_reproduced_and_optional(store):
    codes, queries, idx = store
    q = queries[7]                                      # all ones
    pos, score = idx.binary_search(q, 5, score_mode="popcount")
    assert pos[:2].tolist() == [17, 50_000] and score[0] == 1152      # ties in scan order
    pos_w, score_w = idx.binary_search(q, 5, score_mode="reference")
    assert score_w.max() <= 255
    assert 17 not in pos_w.tolist() or score_w[pos_w.tolist().index(17)] == 1152 % 256


def test_mask_and_tie_sequence(store):
    codes, queries, idx = store
    n = codes.shape[0]
    rng = np.random.default_rng(9)
    admitted = rng.random(n) < 0.4
    perm = rng.permutation(n).astype(np.uint32)         # order[pos] = tie-break sequence of row pos
    seq_sorted = np.argsort(perm)                       # positions in arrival order
    arrival = [int(p) for p in seq_sorted if admitted[p]]
    try:
        idx.set_code_mask(admitted)                     # ties by position
        for k in (20, 500):
            pos, score = idx.binary_search(queries[1], k, use_mask=True)
            opos, oscore = obinary.search(codes, queries[1], k, order=np.flatnonzero(admitted))
            assert np.array_equal(pos, opos) and np.array_equal(score, oscore)
        idx.set_code_mask(admitted, perm)               # ties by the given sequence
        for k in (20, 500):
            pos, score = idx.binary_search(queries[1], k, use_mask=True)
            opos, oscore = obinary.search(codes, queries[1], k, order=arrival)
            assert np.array_equal(pos, opos) and np.array_equal(score, oscore)
        # without use_mask the mask (and its sequence) is ignored
        pos, _ = idx.binary_search(queries[1], 20)
        assert np.array_equal(pos, obinary.search(codes, queries[1], 20)[0])
    finally:
        idx.clear_code_mask()


@pytest.mark.parametrize("n,k", [(1, 1), (1, 5), (255, 20), (256, 20), (257, 300), (70_001, 0), (513, 513)])
def test_small_and_ragged_code_stores(gpu, n, k):
    rng = np.random.default_rng(n)
    codes = (rng.random((n, DIM)) < 0.5).astype(np.uint8)
    q = (rng.random(DIM) < 0.5).astype(np.uint8)
    ids = np.arange(1000, 1000 + n, dtype=np.int64)
    with gpu(0) as idx:
        idx.load_codes(codes, ids)
        assert idx.num_codes == n
        pos, score = idx.binary_search(q, k, score_mode="popcount")
        opos, oscore = obinary.search(codes, q, k, wrap=False)
        assert np.array_equal(pos, opos + 1000) and np.array_equal(score, oscore)


def test_rejects_bytes_that_are_not_sign_codes(gpu):
    from clip_database_b200 import _lib
    codes = np.zeros((40, DIM), dtype=np.uint8)
    codes[3, 700] = 2
    with gpu(0) as idx:
        with pytest.raises(_lib.ClipdbError):
            idx.load_codes(codes)
        assert idx.num_codes == 0
        with pytest.raises(_lib.ClipdbError):
            idx.binary_search(np.zeros(DIM, dtype=np.uint8), 5)      # nothing loaded: an error, no fallback


def test_independent_of_the_float_store(gpu):
    rows = synth.unit_rows(3000, DIM, 5)
    with gpu(0) as idx:
        idx.load(rows)
        idx.load_codes((rows >= 0).astype(np.uint8))
        q = synth.unit_rows(1, DIM, 6)[0]
        before = idx.search(q, 10)
        pos, score = idx.binary_search((q >= 0).astype(np.uint8), 10, score_mode="popcount")
        after = idx.search(q, 10)
        assert np.array_equal(before.rowids, after.rowids)
        assert np.array_equal(pos, obinary.search((rows >= 0).astype(np.uint8), (q >= 0).astype(np.uint8), 10,
                                                  wrap=False)[0])


@pytest.mark.parametrize("name", cpu.case_names())
def test_image_database_reproduces_reference_binary_search(name, tmp_path):
    """The packaged mirror of search() on a binary-only database == what the reference returned."""
    import golden_cases
    from clip_database_b200 import ImageDatabase
    case = next(c for c in cpu.load_golden()["cases"] if c["name"] == name)
    rows, paths, kwargs, vectors = cpu.inputs_for(case)
    e1, e2, weights, negs, ws = golden_cases.embedding_call(kwargs, vectors)
    db_path = str(tmp_path / "b.db")
    synth.write_reference_db(db_path, rows, paths, vectors=False)
    db = ImageDatabase(db_path, device=0)
    try:
        before = db.index.launch_count
        results = db.search_embedding(e1, k=kwargs["k"], embedding2=e2, weights=weights, negative_embeddings=negs,
                                      negative_weights=ws, filter_folders=kwargs.get("filter_folders"),
                                      show_duplicates=kwargs["show_duplicates"])
        assert kwargs["k"] == 0 or db.index.launch_count > before, "the CUDA path did not run"
    finally:
        db.close()
    pos = {p: i for i, p in enumerate(paths)}
    assert [pos[p] for p, _ in results] == case["expected_positions"]
    assert [s for _, s in results] == case["expected_similarities"]
