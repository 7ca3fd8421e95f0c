"""Fused shard exchange (clipdb_search_sharded_device): the scan kernel's last CTA pushes its
shard's top-k into every peer's inbox, waits for theirs and merges — one launch per shard, no
collective.  On a 1-GPU box the ranks are contexts on the SAME device (small grids so the
kernels are co-resident) wired with raw pointers; the multi-process CUDA-IPC wiring over NVLink
is covered by tests/test_gpu_sharded.py on >= 2 GPUs."""
import threading

import numpy as np
import pytest

from clip_database_b200 import synth
from clip_database_b200.sharded import shard_bounds

from conftest import have_gpu

pytestmark = pytest.mark.gpu
DIM = 1152


def run_ranks(torch, shards, fn):
    """fn(rank, idx, stream) on one thread per rank (ctypes releases the GIL)."""
    errs = []

    def work(r):
        try:
            fn(r, shards[r])
        except Exception as e:      # noqa: BLE001
            errs.append((r, e))
    ts = [threading.Thread(target=work, args=(r,)) for r in range(len(shards))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


@pytest.mark.parametrize("world,k", [(2, 20), (3, 100), (4, 1)])
def test_same_device_ranks_equal_unsharded(world, k):
    assert have_gpu()
    import torch
    from clip_database_b200 import GpuIndex
    n = 60_000
    rows = synth.unit_rows(n, DIM, 1234)
    rows[n - 1] = rows[3]                         # a cross-shard exact tie
    rows[20_001] = 0                              # a NaN row in some shard
    queries = synth.unit_rows(5, DIM, 99)
    queries[4] = rows[3]
    with GpuIndex(0) as whole:
        whole.load(rows, np.arange(1, n + 1))
        want = whole.search(queries, k)
    bounds = shard_bounds(n, world)
    shards, inboxes = [], []
    try:
        for r, (lo, hi) in enumerate(bounds):
            s = GpuIndex(0)
            s.load(rows[lo:hi], np.arange(lo + 1, hi + 1))
            s.set_option("scan_ctas", max(148 // world // 2, 1))   # all ranks' kernels co-resident on one GPU
            s.set_option("xchg_timeout_ms", 20000)
            _, ptr = s.exchange_init(world, r)
            shards.append(s)
            inboxes.append(ptr)
        for s in shards:
            s.exchange_connect_pointers(inboxes, [0] * world)
        d_q = torch.from_numpy(queries).cuda()
        outs = [(torch.empty(k, dtype=torch.int64, device="cuda"), torch.empty(k, dtype=torch.float32, device="cuda"),
                 torch.zeros(1, dtype=torch.int32, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"))
                for _ in range(world)]
        torch.cuda.synchronize()
        for qi in range(queries.shape[0]):
            def one(r, s, qi=qi):
                o_ids, o_d, o_n, o_nan = outs[r]
                s.search_sharded_device(d_q[qi], k, o_ids, o_d, o_n, o_nan)
                s.synchronize()
            run_ranks(torch, shards, one)
            for r in range(world):
                o_ids, o_d, o_n, o_nan = outs[r]
                assert int(o_n[0]) == want.counts[qi], "a peer timed out" if int(o_n[0]) < 0 else "count"
                assert np.array_equal(o_ids.cpu().numpy(), want.rowids[qi])
                assert np.array_equal(o_d.cpu().numpy().view(np.uint32), want.distances[qi].view(np.uint32))
                assert int(o_nan[0]) == want.nan_rows[qi]
    finally:
        for s in shards:
            s.close()


def test_missing_peer_times_out_instead_of_hanging():
    assert have_gpu()
    import torch
    from clip_database_b200 import GpuIndex
    rows = synth.unit_rows(5000, DIM, 1)
    with GpuIndex(0) as a, GpuIndex(0) as b:
        a.load(rows[:2500])
        b.load(rows[2500:])
        for r, s in enumerate((a, b)):
            s.set_option("scan_ctas", 32)
            s.set_option("xchg_timeout_ms", 300)
        _, pa = a.exchange_init(2, 0)
        _, pb = b.exchange_init(2, 1)
        a.exchange_connect_pointers([pa, pb], [0, 0])
        b.exchange_connect_pointers([pa, pb], [0, 0])
        q = torch.from_numpy(synth.unit_rows(1, DIM, 2)).cuda()
        o = (torch.empty(5, dtype=torch.int64, device="cuda"), torch.empty(5, dtype=torch.float32, device="cuda"),
             torch.zeros(1, dtype=torch.int32, device="cuda"))
        a.search_sharded_device(q[0], 5, o[0], o[1], o[2])      # rank 1 never searches
        a.synchronize()
        assert int(o[2][0]) == -1


@pytest.mark.parametrize("k", [20, 100])
def test_single_process_multi_gpu_index(k):
    """MultiGpuIndex: the shards are contexts of this process — on every visible GPU when there are
    several (real peer stores over NVLink), else three shards on the one GPU."""
    assert have_gpu()
    import torch
    from clip_database_b200 import GpuIndex
    from clip_database_b200.multigpu import MultiGpuIndex
    n_dev = torch.cuda.device_count()
    devices = list(range(min(n_dev, 4))) if n_dev > 1 else [0, 0, 0]
    n = 50_000
    rows = synth.unit_rows(n, DIM, 4321)
    rows[n - 2] = rows[10]
    rows[30_000] = 0
    queries = synth.unit_rows(4, DIM, 77)
    queries[3] = rows[10]
    with GpuIndex(0) as whole:
        whole.load(rows, np.arange(1, n + 1))
        want = whole.search(queries, k)
        mask = np.arange(n) % 5 != 0
        whole.set_mask(mask)
        want_masked = whole.search(queries, k, use_mask=True)
    with MultiGpuIndex(devices, scan_ctas=None if n_dev > 1 else 24, timeout_ms=20000) as multi:
        multi.load(rows, np.arange(1, n + 1))
        for qi in range(4):
            before = multi.launch_count
            ids, dist, nan = multi.search(queries[qi], k)
            assert multi.launch_count - before == len(devices), "one launch per shard"
            assert np.array_equal(ids, want.rowids[qi]) and nan == want.nan_rows[qi]
            assert np.array_equal(dist.view(np.uint32), want.distances[qi].view(np.uint32))
        multi.set_mask(mask)
        ids, dist, nan = multi.search(queries[3], k, use_mask=True)
        assert np.array_equal(ids, want_masked.rowids[3])
        assert np.array_equal(dist.view(np.uint32), want_masked.distances[3].view(np.uint32))


def test_ranks_asking_different_k_are_detected():
    assert have_gpu()
    import torch
    from clip_database_b200 import GpuIndex
    rows = synth.unit_rows(5000, DIM, 1)
    with GpuIndex(0) as a, GpuIndex(0) as b:
        a.load(rows[:2500])
        b.load(rows[2500:])
        for s in (a, b):
            s.set_option("scan_ctas", 32)
            s.set_option("xchg_timeout_ms", 20000)
        _, pa = a.exchange_init(2, 0)
        _, pb = b.exchange_init(2, 1)
        a.exchange_connect_pointers([pa, pb], [0, 0])
        b.exchange_connect_pointers([pa, pb], [0, 0])
        q = torch.from_numpy(synth.unit_rows(1, DIM, 2)).cuda()
        outs = [(torch.empty(9, dtype=torch.int64, device="cuda"), torch.empty(9, dtype=torch.float32, device="cuda"),
                 torch.zeros(1, dtype=torch.int32, device="cuda")) for _ in range(2)]
        torch.cuda.synchronize()
        a.search_sharded_device(q[0], 5, *outs[0])
        b.search_sharded_device(q[0], 9, *outs[1])
        a.synchronize()
        b.synchronize()
        assert int(outs[0][2][0]) == -2 and int(outs[1][2][0]) == -2


def test_single_process_multi_gpu_batch():
    """MultiGpuIndex.search_batch: bf16 pre-selection on every GPU, candidates exchanged by the batched
    path's last kernel over peer memory.  Needs one GPU per shard (skipped on a 1-GPU box)."""
    assert have_gpu()
    import torch
    from clip_database_b200 import GpuIndex
    from clip_database_b200.multigpu import MultiGpuIndex
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 90_000
    rows = synth.unit_rows(n, DIM, 808)
    rows[n - 3] = rows[20]
    queries = synth.unit_rows(300, DIM, 809)
    queries[11] = 0                                   # flagged on every shard -> re-run through the fused scan
    queries[12] = rows[20]
    with GpuIndex(0) as whole:
        whole.load(rows, np.arange(1, n + 1))
        want = whole.search(queries, 50)
    with MultiGpuIndex(list(range(min(n_dev, 4)))) as multi:
        multi.load(rows, np.arange(1, n + 1))
        multi.enable_batch()
        got = multi.search_batch(queries, 50)
    assert np.array_equal(got.counts, want.counts)
    assert np.array_equal(got.rowids, want.rowids)
    assert np.array_equal(got.distances.view(np.uint32), want.distances.view(np.uint32))
    assert got.rowids[12, :2].tolist() == [21, n - 2]


def test_multi_gpu_index_recovers_when_one_shard_fails_to_launch():
    """a host-side failure on shard r after shards < r have launched used to leave the shards'
    sequence numbers apart for good.  Now refusable requests are refused before the first launch, and a failure
    in the middle resynchronises the exchange: the next search works."""
    assert have_gpu()
    import torch
    from clip_database_b200 import GpuIndex
    from clip_database_b200.multigpu import MultiGpuIndex
    n_dev = torch.cuda.device_count()
    devices = list(range(min(n_dev, 4))) if n_dev > 1 else [0, 0, 0]
    n = 30_000
    rows = synth.unit_rows(n, DIM, 515)
    queries = synth.unit_rows(3, DIM, 516)
    with GpuIndex(0) as whole:
        whole.load(rows, np.arange(1, n + 1))
        want = whole.search(queries, 10)
    with MultiGpuIndex(devices, scan_ctas=None if n_dev > 1 else 24, timeout_ms=3000) as multi:
        multi.load(rows, np.arange(1, n + 1))
        ids, dist, _ = multi.search(queries[0], 10)
        assert np.array_equal(ids, want.rowids[0])
        # refused up front: nothing was launched, nothing to repair
        before = multi.launch_count
        with pytest.raises(ValueError):
            multi.search(queries[1], 10, use_mask=True)                  # no mask installed
        with pytest.raises(ValueError):
            multi.search(queries[1][:100], 10)                          # wrong dimension
        assert multi.launch_count == before
        # a failure in the middle: shard 0 launches, the last shard raises
        victim = multi.shards[-1]
        real = victim.search_sharded_device
        calls = {"n": 0}

        def flaky(*a, **kw):
            calls["n"] += 1
            raise RuntimeError("injected launch failure")
        victim.search_sharded_device = flaky
        with pytest.raises(RuntimeError, match="injected"):
            multi.search(queries[1], 10)
        victim.search_sharded_device = real
        assert calls["n"] == 1
        for qi in range(3):                                              # back in step
            ids, dist, _ = multi.search(queries[qi], 10)
            assert np.array_equal(ids, want.rowids[qi])
            assert np.array_equal(dist.view(np.uint32), want.distances[qi].view(np.uint32))
