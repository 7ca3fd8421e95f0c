"""Parity of the CUDA path (through the C ABI) with the CPU oracle.

Bar (BASELINE.json north_star): top-k row ids identical to the reference's,
except where distances tie within 1e-5 relative; distances within that
tolerance (taken as |delta| <= 1e-5 * max(|d|, 1), see conftest.RTOL).
Selection and merging are integer work on 64-bit keys and are checked bit-exact
against a key-level restatement.
"""
import numpy as np
import pytest

from clip_database_b200 import synth
from oracle import blend as oblend
from oracle import ref

from conftest import assert_topk_parity, have_gpu, tol


def blend_close(got, exp):
    """K3 tolerance.  Every elementwise step is the same IEEE operation as numpy's; only
    the squared-norm summation order differs, so the norm may be off by an ulp and, after
    the `e - w*neg` cancellation, an element by a few ulps OF THE VECTOR'S SCALE:
    |delta| <= 1e-6*|exp_i| + 3e-7*max|exp|  (far inside the 1e-5 of north_star)."""
    exp = np.asarray(exp, dtype=np.float64)
    bound = 1e-6 * np.abs(exp) + 3e-7 * np.abs(exp).max()
    return bool(np.all(np.abs(np.asarray(got, dtype=np.float64) - exp) <= bound))

pytestmark = pytest.mark.gpu

DIM = 1152


@pytest.fixture(scope="module")
def gpu():
    if not have_gpu():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    from clip_database_b200 import GpuIndex
    return GpuIndex


@pytest.fixture(scope="module")
def config1(gpu):
    """Config 1 of BASELINE.json: 100k x 1152 unit rows (seed 1234), 32 queries (seed 99)."""
    rows = synth.unit_rows(100_000, DIM, 1234)
    queries = synth.unit_rows(32, DIM, 99)
    idx = gpu(0)
    idx.load(rows, np.arange(1, rows.shape[0] + 1))
    yield rows, queries, idx
    idx.close()


def check_query(idx, rows, q, k, rowid_offset=1, mask=None, **kw):
    res = idx.search(q, k, **kw)
    ids, d = res.row(0)
    dist_all = ref.distances(rows, q)
    oids, od, oseq, n_nan = ref.knn(rows, q, k, mask=None if mask is None else mask.astype(np.uint8))
    assert res.nan_rows[0] == n_nan
    swaps = assert_topk_parity(ids - rowid_offset, d, oseq, od, dist_all)
    return swaps, res


@pytest.mark.parametrize("variant", [1, 2])
def test_config1_k20_matches_oracle(config1, variant):
    rows, queries, idx = config1
    idx.set_option("scan_variant", variant)
    swaps = 0
    for q in queries:
        s, _ = check_query(idx, rows, q, 20)
        swaps += s
    idx.set_option("scan_variant", 0)
    assert swaps <= 4        # near-tie swaps allowed by the tolerance are rare


@pytest.mark.parametrize("k", [1, 2, 31, 32, 33, 64, 65, 100, 128, 129, 1000])
def test_k_sweep(config1, k):
    rows, queries, idx = config1
    check_query(idx, rows, queries[0], k)
    check_query(idx, rows, queries[1], k)


def test_batched_host_call_equals_one_at_a_time(config1):
    rows, queries, idx = config1
    res = idx.search(queries[:8], 20)
    for i in range(8):
        one = idx.search(queries[i], 20)
        assert np.array_equal(res.rowids[i], one.rowids[0])
        assert np.array_equal(res.distances[i], one.distances[0])


def test_results_are_deterministic(config1):
    rows, queries, idx = config1
    a = idx.search(queries[3], 100)
    for _ in range(3):
        b = idx.search(queries[3], 100)
        assert np.array_equal(a.rowids, b.rowids) and np.array_equal(a.distances, b.distances)


def test_variants_agree_bit_for_bit_on_ids(config1):
    """Both scan variants use the same per-row reduction tree, so keys are identical."""
    rows, queries, idx = config1
    idx.set_option("scan_variant", 1)
    a = idx.search(queries[:4], 64)
    idx.set_option("scan_variant", 2)
    b = idx.search(queries[:4], 64)
    idx.set_option("scan_variant", 0)
    assert np.array_equal(a.rowids, b.rowids)
    assert np.array_equal(a.distances, b.distances)


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("assign", [0, 1, 2])
def test_ring_shapes_and_tile_schedulers_agree(gpu, config1, cfg, assign):
    """Every ring shape / tile-assignment policy of the TMA kernel returns the same bits."""
    rows, queries, idx = config1
    base = idx.search(queries[:3], 20)
    idx.set_option("scan_cfg", cfg)
    idx.set_option("scan_assign", assign)
    idx.set_option("scan_chunk", 3)
    try:
        got = idx.search(queries[:3], 20)
        assert np.array_equal(got.rowids, base.rowids)
        assert np.array_equal(got.distances, base.distances)
        # ragged sizes: fewer tiles than CTAs, partial last tile
        for n in (1, 5, 8, 9, 1183, 1185, 4099):
            with gpu(0) as small:
                small.load(rows[:n])
                want = small.search(queries[0], 20)
                small.set_option("scan_cfg", cfg)
                small.set_option("scan_assign", assign)
                small.set_option("scan_chunk", 3)
                have = small.search(queries[0], 20)
                assert np.array_equal(have.rowids, want.rowids) and np.array_equal(have.counts, want.counts)
                assert np.array_equal(have.distances[:, :want.counts[0]], want.distances[:, :want.counts[0]])
    finally:
        idx.set_option("scan_cfg", 0)
        idx.set_option("scan_assign", 2)
        idx.set_option("scan_chunk", 4)


def test_prefilter_never_changes_the_answer(gpu):
    """Clustered data (many rows within 1e-6 of the k-th distance) must give the same ids as
    the oracle: the float32 pre-filter only skips rows that provably cannot enter."""
    rng = np.random.default_rng(17)
    base = synth.unit_rows(1, DIM, 18)[0]
    rows = base[None, :] + 1e-4 * rng.standard_normal((30_000, DIM), dtype=np.float32)
    rows = (rows / np.linalg.norm(rows, axis=1, keepdims=True)).astype(np.float32)
    q = base + 1e-4 * rng.standard_normal(DIM, dtype=np.float32)
    with gpu(0) as idx:
        idx.load(rows)
        for variant in (1, 2):
            idx.set_option("scan_variant", variant)
            for k in (20, 100):
                ids, d = idx.search(q, k).row(0)
                full_ids, full_d = idx.search(q, rows.shape[0]).row(0)      # general path: no pre-filter
                assert np.array_equal(ids, full_ids[:k])
                assert np.array_equal(d, full_d[:k])


def test_key_level_selection_is_bit_exact(config1):
    """Integer part of the path: with the GPU's own float32 distances for all rows
    (general-k path, k = n), every smaller k must equal the prefix of that full
    ordering exactly — ids and distance bits."""
    rows, queries, idx = config1
    n = rows.shape[0]
    full = idx.search(queries[5], n)
    assert full.counts[0] == n
    ids_full, d_full = full.row(0)
    # the full ordering is sorted by (distance, rowid)
    order = np.lexsort((ids_full, d_full))
    assert np.array_equal(order, np.arange(n))
    for k in (1, 20, 32, 100, 128, 129, 5000):
        part = idx.search(queries[5], k)
        assert np.array_equal(part.rowids[0], ids_full[:k])
        assert np.array_equal(part.distances[0].view(np.uint32), d_full[:k].view(np.uint32))


def test_exact_ties_in_rowid_order(gpu):
    rows = synth.unit_rows(20_000, DIM, 77)
    for dst in (900, 1500, 19_999, 12_345):
        rows[dst] = rows[7]
    with gpu(0) as idx:
        idx.load(rows, np.arange(1, 20_001))
        for variant in (1, 2):
            idx.set_option("scan_variant", variant)
            res = idx.search(rows[7], 5)
            ids, d = res.row(0)
            assert ids.tolist()[:5] == [8, 901, 1501, 12_346, 20_000]
            assert np.all(d[:5] == d[0]) and abs(d[0]) <= 1e-6
            # a later row tying the k-th is not admitted
            assert idx.search(rows[7], 2).rowids[0].tolist() == [8, 901]
            _, od, oseq, _ = ref.knn(rows, rows[7], 5)
            assert oseq.tolist() == [7, 900, 1500, 12_345, 19_999]


@pytest.mark.parametrize("n,k", [(1, 1), (1, 5), (7, 3), (8, 8), (9, 20), (300, 0), (300, 300), (300, 350),
                                 (4097, 128), (4097, 4097)])
def test_small_and_ragged_stores(gpu, n, k):
    rows = synth.unit_rows(n, DIM, 5)
    q = synth.unit_rows(1, DIM, 6)[0]
    with gpu(0) as idx:
        idx.load(rows)                       # implicit rowids = position
        for variant in (1, 2):
            idx.set_option("scan_variant", variant)
            res = idx.search(q, k)
            assert res.counts[0] == min(k, n)
            if k > 0:
                ids, d = res.row(0)
                _, od, oseq, _ = ref.knn(rows, q, k)
                assert_topk_parity(ids, d, oseq, od, ref.distances(rows, q))


def test_nan_rows_are_counted_and_excluded(gpu):
    rows = synth.unit_rows(5000, DIM, 8)
    rows[17] = 0
    rows[4000] = 0
    q = synth.unit_rows(1, DIM, 9)[0]
    with gpu(0) as idx:
        idx.load(rows)
        for variant in (1, 2):
            idx.set_option("scan_variant", variant)
            for k in (20, 200, 5000):
                res = idx.search(q, k)
                ids, d = res.row(0)
                assert res.nan_rows[0] == 2
                assert 17 not in ids and 4000 not in ids
                assert res.counts[0] == min(k, 4998)
                assert not np.isnan(d).any()
        # a zero query makes every distance NaN
        res = idx.search(np.zeros(DIM, dtype=np.float32), 10)
        assert res.counts[0] == 0 and res.nan_rows[0] == 5000


def test_admission_mask(config1):
    rows, queries, idx = config1
    rng = np.random.default_rng(3)
    mask = rng.random(rows.shape[0]) < 0.3
    idx.set_mask(mask)
    for variant in (1, 2):
        idx.set_option("scan_variant", variant)
        for k in (20, 200):
            res = idx.search(queries[2], k, use_mask=True)
            ids, d = res.row(0)
            _, od, oseq, _ = ref.knn(rows, queries[2], k, mask=mask.astype(np.uint8))
            assert_topk_parity(ids - 1, d, oseq, od, ref.distances(rows, queries[2]))
            assert mask[ids - 1].all()
    idx.set_option("scan_variant", 0)
    # an empty admission set
    idx.set_mask(np.zeros(rows.shape[0], dtype=bool))
    assert idx.search(queries[2], 20, use_mask=True).counts[0] == 0
    idx.clear_mask()
    assert idx.search(queries[2], 20).counts[0] == 20


@pytest.mark.parametrize("dim", [4, 64, 100, 1150, 1156, 2048])
def test_other_dimensions_use_the_generic_kernel(gpu, dim):
    rows = synth.unit_rows(3000, dim, 40)
    q = synth.unit_rows(1, dim, 41)[0]
    with gpu(0) as idx:
        idx.load(rows)
        for k in (20, 200):
            res = idx.search(q, k)
            ids, d = res.row(0)
            _, od, oseq, _ = ref.knn(rows, q, k)
            assert_topk_parity(ids, d, oseq, od, ref.distances(rows, q))


def test_l2_metric(config1):
    rows, queries, idx = config1
    q = queries[4]
    exp = np.sqrt(((rows.astype(np.float64) - q.astype(np.float64)) ** 2).sum(axis=1))
    order = np.argsort(exp, kind="stable")[:20]
    for variant in (1, 2):
        idx.set_option("scan_variant", variant)
        ids, d = idx.search(q, 20, metric="l2").row(0)
        assert_topk_parity(ids - 1, d, order, exp[order], exp)
    idx.set_option("scan_variant", 0)


def test_fp16_normalised_inputs(gpu):
    rows = synth.fp16_normalised(synth.unit_rows(30_000, DIM, 50))
    q = synth.fp16_normalised(synth.unit_rows(1, DIM, 51)[0])
    with gpu(0) as idx:
        idx.load(rows)
        ids, d = idx.search(q, 20).row(0)
        _, od, oseq, _ = ref.knn(rows, q, 20)
        assert_topk_parity(ids, d, oseq, od, ref.distances(rows, q))


def test_append_and_update(gpu):
    rows = synth.unit_rows(10_000, DIM, 60)
    q = synth.unit_rows(1, DIM, 61)[0]
    with gpu(0) as idx:
        idx.load(rows[:6000], np.arange(1, 6001))
        idx.append(rows[6000:9000], np.arange(6001, 9001))
        idx.append(rows[9000:], np.arange(9001, 10_001))
        assert idx.num_rows == 10_000
        ids, d = idx.search(q, 50).row(0)
        _, od, oseq, _ = ref.knn(rows, q, 50)
        assert_topk_parity(ids - 1, d, oseq, od, ref.distances(rows, q))
        # UPDATE vec0 SET embedding = ? WHERE rowid = ? (image_database.py:1165-1167)
        idx.update_row(123, q)
        ids, d = idx.search(q, 1).row(0)
        assert ids[0] == 124 and abs(d[0]) < 1e-6


def test_attach_borrows_device_memory(gpu):
    import torch
    rows = synth.unit_rows(50_000, DIM, 70)
    q = synth.unit_rows(1, DIM, 71)[0]
    t = torch.from_numpy(rows).cuda()
    with gpu(0) as idx:
        idx.attach(t, rowid_base=1000)
        ids, d = idx.search(q, 20).row(0)
        _, od, oseq, _ = ref.knn(rows, q, 20)
        assert_topk_parity(ids - 1000, d, oseq, od, ref.distances(rows, q))
        # device-pointer entry point on torch's stream
        idx.use_torch_stream()
        dq = torch.from_numpy(q).cuda()
        o_id = torch.empty((1, 20), dtype=torch.int64, device="cuda")
        o_d = torch.empty((1, 20), dtype=torch.float32, device="cuda")
        o_n = torch.empty(1, dtype=torch.int32, device="cuda")
        o_nan = torch.empty(1, dtype=torch.int64, device="cuda")
        idx.search_device(dq.view(1, -1), 20, o_id, o_d, o_n, o_nan)
        torch.cuda.synchronize()
        assert int(o_n[0]) == 20 and int(o_nan[0]) == 0
        assert np.array_equal(o_id.cpu().numpy()[0], ids)
        assert np.array_equal(o_d.cpu().numpy()[0], d)
        idx.set_stream(None)


# ---- K3: blend / negatives -------------------------------------------------------------

def blend_cases():
    e1 = synth.unit_rows(1, DIM, 201)[0]
    e2 = synth.unit_rows(1, DIM, 202)[0]
    n1 = synth.unit_rows(1, DIM, 203)[0]
    n2 = synth.unit_rows(1, DIM, 204)[0]
    n3 = synth.unit_rows(1, DIM, 205)[0]
    sp1 = np.zeros(DIM, dtype=np.float32); sp1[0:4] = 0.5
    sp2 = np.zeros(DIM, dtype=np.float32); sp2[10:14] = 0.5
    spb = (np.float32(0.5) * sp1 + np.float32(0.5) * sp2) / np.sqrt(np.float32(0.5))
    return {
        "single": (e1, None, (0.5, 0.5), [], [], 0),
        "blend_07_03": (e1, e2, (0.7, 0.3), [], [], 0),
        "blend_raw_weights": (e1, e2, (2.0, 6.0), [], [], 0),
        "blend_zero_weights": (e1, e2, (0.0, 0.0), [], [], 0),
        "blend_negative": (e1, e2, (0.7, 0.3), [n1], [0.5], 0),
        "three_negatives": (e1, None, (0.5, 0.5), [n1, n2, n3], [0.5, 0.25, 1.5], 0),
        "blend_cancels": (e1, -e1, (0.5, 0.5), [], [], 1),
        "negative_cancels": (e1, None, (0.5, 0.5), [e1], [1.0], 2),
        "negative_cancels_reblend": (sp1, sp2, (0.5, 0.5), [spb], [1.0], 2),
        "unnormalised_input": (3.0 * e1, 0.5 * e2, (0.6, 0.4), [2.0 * n1], [0.1], 0),
    }


@pytest.mark.parametrize("name", list(blend_cases()))
def test_blend_matches_numpy_restatement(gpu, name):
    e1, e2, w, negs, nws, want_flags = blend_cases()[name]
    exp = oblend.compose_query(e1, e2, w, negs, nws)
    with gpu(0) as idx:
        got, flags = idx.blend(e1, e2, w, negs, nws)
    assert flags == want_flags
    # float tolerance: only the squared-norm summation order differs from numpy's BLAS dot
    assert blend_close(got, exp), np.abs(got - exp).max()
    if name in ("single", "blend_cancels", "negative_cancels"):
        assert np.array_equal(got, exp)      # pure copies must be bit-exact


def test_blend_matches_reference_outputs(gpu, golden):
    by = {b["name"]: np.array(b["out"], dtype=np.float32) for b in golden["blend_cases"]}
    e1 = synth.unit_rows(1, DIM, 201)[0]
    e2 = synth.unit_rows(1, DIM, 202)[0]
    n1 = synth.unit_rows(1, DIM, 203)[0]
    n2 = synth.unit_rows(1, DIM, 204)[0]
    with gpu(0) as idx:
        assert blend_close(idx.blend(e1, None, (0.5, 0.5), [n1], [0.5])[0], by["one_negative"])
        assert blend_close(idx.blend(e1, None, (0.5, 0.5), [n1, n2], [0.5, 0.8])[0], by["two_negatives"])
        assert np.array_equal(idx.blend(e1, None, (0.5, 0.5), [e1], [1.0])[0], by["zero_restores_e1"])


def test_blend_search_equals_blend_then_search(config1):
    rows, queries, idx = config1
    e1, e2, neg = queries[10], queries[11], queries[12]
    res, q_gpu, flags = idx.blend_search(e1, 20, e2=e2, weights=(0.7, 0.3), negatives=[neg],
                                         negative_weights=[0.5], return_query=True)
    q_sep, _ = idx.blend(e1, e2, (0.7, 0.3), [neg], [0.5])
    assert np.array_equal(q_gpu, q_sep) and flags == 0
    two = idx.search(q_sep, 20)
    assert np.array_equal(res.rowids, two.rowids) and np.array_equal(res.distances, two.distances)
    # and against the oracle, end to end (config 4 of BASELINE.json)
    q_ref = oblend.compose_query(e1, e2, (0.7, 0.3), [neg], [0.5])
    _, od, oseq, _ = ref.knn(rows, q_ref, 20)
    assert_topk_parity(res.rowids[0] - 1, res.distances[0], oseq, od, ref.distances(rows, q_ref))


def test_batched_blend_device(gpu):
    import torch
    B, n_neg = 16, 2
    e1 = synth.unit_rows(B, DIM, 301)
    e2 = synth.unit_rows(B, DIM, 302)
    negs = synth.unit_rows(B * n_neg, DIM, 303).reshape(B, n_neg, DIM)
    w = np.tile(np.array([0.7, 0.3], dtype=np.float32), (B, 1))
    nw = np.tile(np.array([0.5, 0.25], dtype=np.float32), (B, 1))
    with gpu(0) as idx:
        idx.use_torch_stream()
        out = torch.empty((B, DIM), dtype=torch.float32, device="cuda")
        flags = torch.empty(B, dtype=torch.int32, device="cuda")
        idx.blend_device(torch.from_numpy(e1).cuda(), torch.from_numpy(e2).cuda(), torch.from_numpy(w).cuda(),
                         torch.from_numpy(negs).cuda(), torch.from_numpy(nw).cuda(), out, flags)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        idx.set_stream(None)
    for b in range(B):
        exp = oblend.compose_query(e1[b], e2[b], (0.7, 0.3), list(negs[b]), [0.5, 0.25])
        assert blend_close(got[b], exp)
    assert int(flags.sum()) == 0


# ---- shard merge --------------------------------------------------------------------------

def test_merge_device_matches_lexicographic_merge(gpu):
    import torch
    rng = np.random.default_rng(11)
    lists, k = 8, 100
    dist = np.sort(rng.random((lists, k)).astype(np.float32), axis=1)
    dist[3, :5] = dist[1, :5]                    # cross-shard exact ties
    dist[3] = np.sort(dist[3]); dist[1] = np.sort(dist[1])
    rowids = (np.arange(lists)[:, None] * 1_000_000 + np.sort(rng.choice(1_000_000, (lists, k)), axis=1)).astype(np.int64)
    counts = np.array([k, k, 37, k, 0, k, 1, k], dtype=np.int32)
    keys = []
    for l in range(lists):
        for p in range(counts[l]):
            keys.append((dist[l, p], l, p))
    keys.sort()
    exp = keys[:k]
    with gpu(0) as idx:
        idx.use_torch_stream()
        o_d = torch.empty(k, dtype=torch.float32, device="cuda")
        o_i = torch.empty(k, dtype=torch.int64, device="cuda")
        o_n = torch.empty(1, dtype=torch.int32, device="cuda")
        idx.merge_device(torch.from_numpy(dist).cuda(), torch.from_numpy(rowids).cuda(),
                         torch.from_numpy(counts).cuda(), k, o_d, o_i, o_n)
        torch.cuda.synchronize()
        idx.set_stream(None)
    assert int(o_n[0]) == k
    assert np.array_equal(o_d.cpu().numpy(), np.array([e[0] for e in exp], dtype=np.float32))
    assert np.array_equal(o_i.cpu().numpy(), np.array([rowids[e[1], e[2]] for e in exp]))


def test_sharded_search_equals_unsharded(config1):
    """Row-sharding logic on one GPU: split the store into G contiguous ranges, search
    each, merge — must equal the single-store answer bit for bit."""
    import torch
    from clip_database_b200 import GpuIndex
    rows, queries, idx = config1
    k, G = 20, 4
    n = rows.shape[0]
    full = idx.search(queries[:4], k)
    bounds = [n * g // G for g in range(G + 1)]
    shards = []
    for g in range(G):
        s = GpuIndex(0)
        s.load(rows[bounds[g]:bounds[g + 1]], np.arange(bounds[g] + 1, bounds[g + 1] + 1))
        shards.append(s)
    for qi in range(4):
        parts = [s.search(queries[qi], k) for s in shards]
        dist = torch.from_numpy(np.stack([p.distances[0] for p in parts])).cuda()
        ids = torch.from_numpy(np.stack([p.rowids[0] for p in parts])).cuda()
        cnt = torch.from_numpy(np.array([p.counts[0] for p in parts], dtype=np.int32)).cuda()
        o_d = torch.empty(k, dtype=torch.float32, device="cuda")
        o_i = torch.empty(k, dtype=torch.int64, device="cuda")
        o_n = torch.empty(1, dtype=torch.int32, device="cuda")
        idx.use_torch_stream()
        idx.merge_device(dist, ids, cnt, k, o_d, o_i, o_n)
        torch.cuda.synchronize()
        idx.set_stream(None)
        assert np.array_equal(o_i.cpu().numpy(), full.rowids[qi])
        assert np.array_equal(o_d.cpu().numpy(), full.distances[qi])
    for s in shards:
        s.close()


def test_adversarial_row_order_every_row_is_admitted(gpu):
    """Rows stored in order of DECREASING distance to the query: every row beats the current
    k-th candidate of its warp (the admission path runs for every row instead of ~k/n of them).
    Still exact, for both scan kernels."""
    rows = synth.unit_rows(60_000, DIM, 606)
    q = synth.unit_rows(1, DIM, 607)[0]
    order = np.argsort(-ref.distances(rows, q), kind="stable")       # farthest first
    rows = np.ascontiguousarray(rows[order[::1]])
    with gpu(0) as idx:
        idx.load(rows)
        for variant in (1, 2):
            idx.set_option("scan_variant", variant)
            for k in (20, 128):
                check_query(idx, rows, q, k, rowid_offset=0)
        got = idx.search(q, 5)
        assert got.rowids[0].tolist() == [59_999, 59_998, 59_997, 59_996, 59_995]


def test_one_context_from_many_threads(config1):
    """Calls on one context are serialised by the library: concurrent host threads get the
    same answers as sequential calls."""
    import threading
    rows, queries, idx = config1
    want = [idx.search(queries[i], 20) for i in range(8)]
    got = [None] * 8
    errs = []

    def work(i):
        try:
            for _ in range(5):
                got[i] = idx.search(queries[i], 20)
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i in range(8):
        assert np.array_equal(got[i].rowids, want[i].rowids)
        assert np.array_equal(got[i].distances.view(np.uint32), want[i].distances.view(np.uint32))


def test_clustered_store_and_queries_near_stored_rows(gpu):
    """SURVEY §8d's clustered variant at config-1 scale: 64 clusters (cosine ~0.64 between mates),
    queries = stored row + 10 % noise, so the top-k is a dense neighbourhood rather than the tail of
    a noise distribution.  Oracle parity for the scan, bit equality for the batched path."""
    rng = np.random.default_rng(2024)
    n, clusters = 100_000, 64
    centres = synth.unit_rows(clusters, DIM, 5150)
    rows = centres[np.arange(n) % clusters] + 0.75 * rng.standard_normal((n, DIM), dtype=np.float32) / np.sqrt(DIM)
    rows = (rows / np.linalg.norm(rows, axis=1, keepdims=True)).astype(np.float32)
    picks = rng.choice(n, 12, replace=False)
    queries = rows[picks] + 0.1 * synth.unit_rows(12, DIM, 77)
    queries = (queries / np.linalg.norm(queries, axis=1, keepdims=True)).astype(np.float32)
    with gpu(0) as idx:
        idx.load(rows)
        for q in queries[:6]:
            for k in (20, 100):
                check_query(idx, rows, q, k, rowid_offset=0)
        exact = idx.search(queries, 100)
        assert np.array_equal(exact.rowids[:, 0], picks)             # the perturbed row itself comes first
        idx.enable_batch()
        got = idx.search(queries, 100)
        assert np.array_equal(got.rowids, exact.rowids)
        assert np.array_equal(got.distances.view(np.uint32), exact.distances.view(np.uint32))
        cand, surv = idx.batch_stats()
        assert cand[:12].max() < idx.get_option("batch_cand_cap")
