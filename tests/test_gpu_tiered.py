"""Stores filled by appends (clipdb_reserve_rows) and TIERED stores: float32 rows split between HBM and pinned,
device-mapped host memory, the bf16 copy resident (a "bf16-primary" store: BASELINE configs[4] on 2 GPUs, where
a shard is 230 GB of float32).  Whatever the placement, every search must return exactly what the all-HBM store
returns: the same kernels read the same bits, only the addresses differ."""
import threading
import time

import numpy as np
import pytest

from clip_database_b200 import synth

from conftest import have_gpu

pytestmark = pytest.mark.gpu
DIM = 1152


def same(a, b):
    assert np.array_equal(a.counts, b.counts)
    assert np.array_equal(a.rowids, b.rowids)
    assert np.array_equal(a.distances.view(np.uint32), b.distances.view(np.uint32))
    assert np.array_equal(a.nan_rows, b.nan_rows)


@pytest.fixture(scope="module")
def data():
    assert have_gpu(), "GPU tests selected but no CUDA device is visible"
    n = 30_000
    rows = synth.unit_rows(n, DIM, 2024)
    rows[29_999] = rows[5]                     # an exact tie across the tier boundary
    rows[17_000] = 0                           # a NaN row in the host tier
    queries = synth.unit_rows(70, DIM, 2025)
    queries[3] = rows[5]
    ids = np.arange(11, n + 11, dtype=np.int64)
    from clip_database_b200 import GpuIndex
    with GpuIndex(0) as whole:
        whole.load(rows, ids)
        want20 = whole.search(queries, 20)
        want100 = whole.search(queries[:8], 100)
        want500 = whole.search(queries[:2], 500)       # radix-sort path (k > 128)
    return rows, ids, queries, want20, want100, want500


@pytest.mark.parametrize("device_rows", [12_800, 0, 29_900])
def test_tiered_store_answers_like_the_resident_one(data, device_rows):
    from clip_database_b200 import GpuIndex
    rows, ids, queries, want20, want100, want500 = data
    with GpuIndex(0) as idx:
        idx.reserve(len(rows), DIM, explicit_rowids=True, placement="host", device_rows=device_rows)
        for lo in range(0, len(rows), 7_000):                       # chunks straddle the tier boundary
            idx.append(rows[lo:lo + 7_000], ids[lo:lo + 7_000])
        assert idx.num_rows == len(rows)
        same(idx.search(queries, 20), want20)                        # exact scan, TMA ring
        same(idx.search(queries[:8], 100), want100)
        same(idx.search(queries[:2], 500), want500)
        idx.set_option("scan_variant", 2)                            # exact scan, direct loads
        same(idx.search(queries[:8], 20), type(want20)(want20.rowids[:8], want20.distances[:8], want20.counts[:8],
                                                       want20.nan_rows[:8]))
        idx.set_option("scan_variant", 0)
        # bf16-primary: pre-selection on the resident bf16 copy, exact re-rank reads the tiers
        idx.enable_batch()
        idx.set_option("batch_min_nq", 1)
        before = idx.launch_count
        got = idx.search(queries, 20)
        assert idx.launch_count - before <= 16, "the batched path was not taken"
        same(got, want20)
        one = idx.search(queries[3], 20)                             # a single query through the pre-selection
        assert np.array_equal(one.rowids[0], want20.rowids[3])
        assert one.rowids[0, :2].tolist() == [16, 30_010]            # the tie: rowid order
        same(idx.search(queries[:8], 100), want100)
        # an in-place update of a host-tier row reaches both the float32 tier and the bf16 copy
        idx.update_row(25_000, queries[9])
        hit = idx.search(queries[9], 3)
        assert hit.rowids[0, 0] == ids[25_000] and abs(hit.distances[0, 0]) < 1e-6


def test_appends_keep_the_bf16_copy_current(data):
    from clip_database_b200 import GpuIndex
    rows, ids, queries, want20, want100, _ = data
    with GpuIndex(0) as idx:
        idx.load(rows[:20_000], ids[:20_000])
        idx.enable_batch()
        first = idx.search(queries[:16], 20)
        idx.append(rows[20_000:26_000], ids[20_000:26_000])          # grows the store and the bf16 copy
        idx.append(rows[26_000:], ids[26_000:])
        before = idx.launch_count
        got = idx.search(queries, 20)
        assert idx.launch_count - before <= 16
        same(got, want20)
        assert not np.array_equal(first.rowids, want20.rowids[:16])  # the appended rows do matter


def test_streamed_load_through_the_stage_buffer(data):
    from clip_database_b200 import GpuIndex
    rows, ids, queries, want20, _, _ = data
    with GpuIndex(0) as idx:
        idx.reserve(1_000, DIM, explicit_rowids=True)               # too small on purpose: appends grow it
        stage = idx.stage_buffer(4_096, DIM)
        assert stage.shape == (4_096, DIM) and stage.dtype == np.float32
        for lo in range(0, len(rows), 4_096):
            m = min(4_096, len(rows) - lo)
            stage[:m] = rows[lo:lo + m]
            idx.append(stage[:m], ids[lo:lo + m])
        same(idx.search(queries[:10], 20), type(want20)(want20.rowids[:10], want20.distances[:10], want20.counts[:10],
                                                        want20.nan_rows[:10]))


def _two_ranks(rows, timeout_ms):
    import torch  # noqa: F401
    from clip_database_b200 import GpuIndex
    a, b = GpuIndex(0), GpuIndex(0)
    half = len(rows) // 2
    a.load(rows[:half], np.arange(1, half + 1))
    b.load(rows[half:], np.arange(half + 1, len(rows) + 1))
    for s in (a, b):
        s.set_option("scan_ctas", 32)
        s.set_option("xchg_timeout_ms", timeout_ms)
    _, pa = a.exchange_init(2, 0)
    _, pb = b.exchange_init(2, 1)
    a.exchange_connect_pointers([pa, pb], [0, 0])
    b.exchange_connect_pointers([pa, pb], [0, 0])
    return a, b


def test_host_abort_ends_a_stuck_exchange():
    """A rank whose peer never shows up spins in-kernel until xchg_timeout_ms; clipdb_exchange_abort ends the
    wait from the host at once."""
    assert have_gpu()
    import torch
    rows = synth.unit_rows(5_000, DIM, 1)
    a, b = _two_ranks(rows, 60_000)
    try:
        q = torch.from_numpy(synth.unit_rows(1, DIM, 2)).cuda()
        o = (torch.empty(5, dtype=torch.int64, device="cuda"), torch.empty(5, dtype=torch.float32, device="cuda"),
             torch.zeros(1, dtype=torch.int32, device="cuda"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.search_sharded_device(q[0], 5, o[0], o[1], o[2])          # rank 1 never searches
        threading.Timer(0.3, a.exchange_abort).start()
        a.synchronize()
        waited = time.perf_counter() - t0
        assert int(o[2][0]) == -1
        assert waited < 10.0, f"abort did not end the wait ({waited:.1f} s)"
        # re-arm, bring the two ranks back in step, and the next search works
        a.exchange_abort(False)
        a.exchange_set_epoch(100, 100)
        b.exchange_set_epoch(100, 100)
        outs = [(torch.empty(5, dtype=torch.int64, device="cuda"), torch.empty(5, dtype=torch.float32, device="cuda"),
                 torch.zeros(1, dtype=torch.int32, device="cuda")) for _ in range(2)]
        a.search_sharded_device(q[0], 5, *outs[0])
        b.search_sharded_device(q[0], 5, *outs[1])
        a.synchronize()
        b.synchronize()
        assert int(outs[0][2][0]) == 5 and int(outs[1][2][0]) == 5
        assert np.array_equal(outs[0][0].cpu().numpy(), outs[1][0].cpu().numpy())
    finally:
        a.close()
        b.close()


def test_exchange_epoch_wraps_around():
    """The sequence number is a uint32 that skips 0: searches across the wrap still pair up."""
    assert have_gpu()
    import torch
    from clip_database_b200 import GpuIndex
    rows = synth.unit_rows(6_000, DIM, 3)
    queries = synth.unit_rows(6, DIM, 4)
    with GpuIndex(0) as whole:
        whole.load(rows, np.arange(1, len(rows) + 1))
        want = whole.search(queries, 10)
    a, b = _two_ranks(rows, 20_000)
    try:
        a.exchange_set_epoch(0xFFFFFFFD, 0)
        b.exchange_set_epoch(0xFFFFFFFD, 0)
        d_q = torch.from_numpy(queries).cuda()
        for qi in range(6):                                          # epochs ...FE, ...FF, (0 skipped) 1, 2, ...
            outs = [(torch.empty(10, dtype=torch.int64, device="cuda"), torch.empty(10, dtype=torch.float32, device="cuda"),
                     torch.zeros(1, dtype=torch.int32, device="cuda")) for _ in range(2)]
            a.search_sharded_device(d_q[qi], 10, *outs[0])
            b.search_sharded_device(d_q[qi], 10, *outs[1])
            a.synchronize()
            b.synchronize()
            for o in outs:
                assert int(o[2][0]) == 10
                assert np.array_equal(o[0].cpu().numpy(), want.rowids[qi])
                assert np.array_equal(o[1].cpu().numpy().view(np.uint32), want.distances[qi].view(np.uint32))
    finally:
        a.close()
        b.close()


def test_exchange_timeline_statistics():
    assert have_gpu()
    import torch
    rows = synth.unit_rows(40_000, DIM, 5)
    a, b = _two_ranks(rows, 20_000)
    try:
        a.exchange_stats(enable=True, reset=True)
        b.exchange_stats(enable=True, reset=True)
        d_q = torch.from_numpy(synth.unit_rows(8, DIM, 6)).cuda()
        outs = [(torch.empty(10, dtype=torch.int64, device="cuda"), torch.empty(10, dtype=torch.float32, device="cuda"),
                 torch.zeros(1, dtype=torch.int32, device="cuda")) for _ in range(2)]
        for qi in range(8):
            a.search_sharded_device(d_q[qi], 10, *outs[0])
            b.search_sharded_device(d_q[qi], 10, *outs[1])
        a.synchronize()
        b.synchronize()
        for s in (a, b):
            stats, launches = s.exchange_stats(enable=False, reset=True)
            assert launches == 8
            assert set(stats) == {"scan", "local_merge", "publish", "wait_peers", "final_merge"}
            assert all(0 <= v < 5e6 for v in stats.values()), stats
            assert stats["scan"] > 0
        stats, launches = a.exchange_stats(enable=False, reset=False)
        assert launches == 0
    finally:
        a.close()
        b.close()
