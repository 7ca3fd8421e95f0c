"""K4: the tensor-core batched path must return exactly what the single-query path returns."""
import numpy as np
import pytest

from clip_database_b200 import synth

from conftest import have_gpu

pytestmark = pytest.mark.gpu
DIM = 1152


@pytest.fixture(scope="module")
def store():
    assert have_gpu()
    from clip_database_b200 import GpuIndex
    rows = synth.unit_rows(100_000, DIM, 1234)
    rows[5] = 0                                  # a zero row: NaN distance for every query
    exact = GpuIndex(0)
    exact.load(rows, np.arange(1, rows.shape[0] + 1))
    batched = GpuIndex(0)
    batched.load(rows, np.arange(1, rows.shape[0] + 1))
    batched.enable_batch()
    yield rows, exact, batched
    exact.close()
    batched.close()


def same(a, b):
    assert np.array_equal(a.counts, b.counts)
    assert np.array_equal(a.nan_rows, b.nan_rows)
    assert np.array_equal(a.rowids, b.rowids)
    assert np.array_equal(a.distances.view(np.uint32), b.distances.view(np.uint32))


@pytest.mark.parametrize("nq,k", [(256, 100), (256, 20), (16, 100), (100, 1), (300, 64), (37, 128), (2, 20)])
def test_batched_equals_exact(store, nq, k):
    rows, exact, batched = store
    queries = synth.unit_rows(nq, DIM, 99)
    before = batched.launch_count
    got = batched.search(queries, k)
    launches = batched.launch_count - before
    assert launches <= 8 * ((nq + 255) // 256) + 8, "batched path was not taken"
    same(got, exact.search(queries, k))


def test_single_cta_contraction_kernel_agrees(store):
    """batch_cta_pair=0 selects the 1-CTA tcgen05 kernel; same answers as the CTA-pair one."""
    rows, exact, batched = store
    queries = synth.unit_rows(256, DIM, 123)
    batched.set_option("batch_cta_pair", 0)
    try:
        same(batched.search(queries, 100), exact.search(queries, 100))
    finally:
        batched.set_option("batch_cta_pair", 1)


def test_batched_clustered_queries(store):
    """Queries near stored rows: top results are far from the noise tail."""
    rows, exact, batched = store
    rng = np.random.default_rng(7)
    picks = rng.choice(rows.shape[0], 64, replace=False)
    picks = picks[picks != 5]
    q = rows[picks] + 0.1 * rng.standard_normal((len(picks), DIM), dtype=np.float32) / np.sqrt(DIM)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    got = batched.search(q, 100)
    same(got, exact.search(q, 100))
    assert np.array_equal(got.rowids[:, 0], picks + 1)


@pytest.mark.parametrize("stride,refine", [(1, 1), (4, 0), (16, 1), (64, 1), (64, 0)])
def test_sampling_stride_and_refinement_do_not_change_results(store, stride, refine):
    """Pass A's sampling density and the second (candidate-derived) threshold only change how
    many rows are re-ranked, never the answer."""
    rows, exact, batched = store
    queries = synth.unit_rows(96, DIM, 77)
    batched.set_option("batch_sample_stride", stride)
    batched.set_option("batch_refine", refine)
    try:
        same(batched.search(queries, 100), exact.search(queries, 100))
        same(batched.search(queries[:40], 7), exact.search(queries[:40], 7))
    finally:
        batched.set_option("batch_sample_stride", 0)
        batched.set_option("batch_refine", 1)


@pytest.mark.parametrize("npass", [64, 128, 256])
def test_queries_per_pass_variants_agree(store, npass):
    """The contraction is instantiated for 64 / 128 / 256 queries per pass (UMMA N); a small batch
    forced through a wider instantiation gives the same answer."""
    rows, exact, batched = store
    queries = synth.unit_rows(40, DIM, 321)
    batched.set_option("batch_npass", npass)
    try:
        same(batched.search(queries, 100), exact.search(queries, 100))
    finally:
        batched.set_option("batch_npass", 0)


def test_single_query_through_the_batched_path(store):
    """batch_min_nq = 1: even one query is pre-selected on the bf16 store (half the bytes of the
    float32 scan) and re-ranked exactly."""
    rows, exact, batched = store
    queries = synth.unit_rows(6, DIM, 55)
    batched.set_option("batch_min_nq", 1)
    try:
        for q in queries:
            before = batched.launch_count
            got = batched.search(q, 20)
            assert batched.launch_count - before == 7, "batched path was not taken"
            same(got, exact.search(q, 20))
    finally:
        batched.set_option("batch_min_nq", 2)


@pytest.mark.parametrize("n,k", [(1, 1), (100, 20), (129, 100), (5000, 128), (40_000, 100)])
def test_small_stores(n, k):
    """Stores smaller than a few tiles: fewer sampled groups than k means every row is a
    candidate; more rows than the candidate capacity means the exact scan takes over."""
    from clip_database_b200 import GpuIndex
    rows = synth.unit_rows(n, DIM, n)
    queries = synth.unit_rows(9, DIM, n + 1)
    with GpuIndex(0) as exact, GpuIndex(0) as batched:
        exact.load(rows)
        batched.load(rows)
        batched.enable_batch()
        kk = min(k, n)
        same(batched.search(queries, kk), exact.search(queries, kk))


def test_batched_with_folder_mask():
    """The folder pre-filter bitset (idb:1509-1530) is honoured by the tensor-core path."""
    from clip_database_b200 import GpuIndex
    rows = synth.unit_rows(80_000, DIM, 21)
    rng = np.random.default_rng(5)
    admitted = rng.random(rows.shape[0]) < 0.3
    admitted[:100] = False
    queries = synth.unit_rows(64, DIM, 22)
    queries[3] = rows[50]            # best match is masked out
    queries[4] = rows[np.flatnonzero(admitted)[10]]
    with GpuIndex(0) as exact, GpuIndex(0) as batched:
        exact.load(rows)
        batched.load(rows)
        exact.set_mask(admitted)
        batched.set_mask(admitted)
        batched.enable_batch()
        before = batched.launch_count
        got = batched.search(queries, 50, use_mask=True)
        assert batched.launch_count - before <= 16, "batched path was not taken"
        same(got, exact.search(queries, 50, use_mask=True))
        assert admitted[got.rowids].all()
        # a mask admitting fewer rows than k: every admitted row comes back
        few = np.zeros(rows.shape[0], dtype=bool)
        few[[7, 70_000, 123]] = True
        exact.set_mask(few)
        batched.set_mask(few)
        got = batched.search(queries[:8], 20, use_mask=True)
        same(got, exact.search(queries[:8], 20, use_mask=True))
        assert np.all(got.counts == 3)


def test_batched_sees_appended_and_updated_rows():
    """append_rows / update_row after enable_batch: the bf16 copy is rebuilt before the next
    batched search (the scanner INSERTs and UPDATEs in place, idb:1165-1175)."""
    from clip_database_b200 import GpuIndex
    rows = synth.unit_rows(70_000, DIM, 31)
    extra = synth.unit_rows(300, DIM, 32)
    queries = synth.unit_rows(16, DIM, 33)
    queries[0] = extra[17]
    queries[1] = extra[200]
    with GpuIndex(0) as exact, GpuIndex(0) as batched:
        exact.load(rows)
        batched.load(rows)
        batched.enable_batch()
        same(batched.search(queries, 10), exact.search(queries, 10))
        exact.append(extra)
        batched.append(extra)
        got = batched.search(queries, 10)
        same(got, exact.search(queries, 10))
        assert got.rowids[0, 0] == 70_000 + 17
        exact.update_row(123, extra[200])
        batched.update_row(123, extra[200])
        got = batched.search(queries, 10)
        same(got, exact.search(queries, 10))
        assert got.rowids[1, 0] == 123 and got.rowids[1, 1] == 70_000 + 200


def test_batched_zero_query_falls_back(store):
    rows, exact, batched = store
    queries = synth.unit_rows(32, DIM, 5)
    queries[7] = 0
    got = batched.search(queries, 20)
    same(got, exact.search(queries, 20))
    assert got.counts[7] == 0 and got.nan_rows[7] == rows.shape[0]


def test_batched_overflow_falls_back():
    """Dense cluster + tiny candidate capacity: the filter overflows, the host re-runs the
    flagged queries through the exact scan, and the answer is still exact."""
    from clip_database_b200 import GpuIndex
    rng = np.random.default_rng(3)
    base = synth.unit_rows(1, DIM, 8)[0]
    rows = base[None, :] + 0.02 * rng.standard_normal((70_001, DIM), dtype=np.float32) / np.sqrt(DIM)
    rows = (rows / np.linalg.norm(rows, axis=1, keepdims=True)).astype(np.float32)
    queries = rows[:24] + 0.001 * rng.standard_normal((24, DIM), dtype=np.float32)
    with GpuIndex(0) as exact, GpuIndex(0) as batched:
        exact.load(rows)
        batched.load(rows)
        batched.set_option("batch_cand_cap", 512)
        batched.enable_batch()
        same(batched.search(queries, 50), exact.search(queries, 50))


def test_candidate_capacity_can_change_after_enable(store):
    """batch_cand_cap sizes device buffers: changing it on a live batch store rebuilds them."""
    rows, exact, batched = store
    queries = synth.unit_rows(50, DIM, 909)
    want = exact.search(queries, 100)
    try:
        for cap in (131072, 256, 32768):     # 256: every query overflows and is re-run exactly
            batched.set_option("batch_cand_cap", cap)
            same(batched.search(queries, 100), want)
    finally:
        batched.set_option("batch_cand_cap", 65536)


def test_batch_device_entry_point(store):
    import torch
    rows, exact, batched = store
    queries = synth.unit_rows(256, DIM, 42)
    k = 100
    dq = torch.from_numpy(queries).cuda()
    o_id = torch.empty((256, k), dtype=torch.int64, device="cuda")
    o_d = torch.empty((256, k), dtype=torch.float32, device="cuda")
    o_n = torch.empty(256, dtype=torch.int32, device="cuda")
    o_nan = torch.empty(256, dtype=torch.int64, device="cuda")
    flags = torch.empty(256, dtype=torch.int32, device="cuda")
    batched.use_torch_stream()
    batched.search_batch_device(dq, k, o_id, o_d, o_n, o_nan, flags)
    torch.cuda.synchronize()
    batched.set_stream(None)
    assert int(flags.abs().sum()) == 0
    want = exact.search(queries, k)
    assert np.array_equal(o_id.cpu().numpy(), want.rowids)
    assert np.array_equal(o_d.cpu().numpy().view(np.uint32), want.distances.view(np.uint32))
    assert np.array_equal(o_nan.cpu().numpy(), want.nan_rows)
