"""Multi-rank orchestration (record layout, one all-gather, merge order) on CPU:
world_size 2 and 3 over gloo, per-rank work done by the oracle standing in for the
CUDA backend.  The sharded answer must equal the unsharded oracle answer exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clip_database_b200 import synth
from clip_database_b200.sharded import RecordLayout, ShardedIndex, shard_bounds
from oracle import ref

DIM = 1152


class OracleShardBackend:
    """CPU stand-in with the same interface as CudaShardBackend (tests only)."""

    def __init__(self, rows, rowid_lo):
        self.rows = rows
        self.rowids = np.arange(rowid_lo, rowid_lo + rows.shape[0], dtype=np.int64)

    def new_buffer(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8)

    def to_device(self, query):
        return torch.from_numpy(np.ascontiguousarray(query, dtype=np.float32))

    def new_outputs(self, k):
        return (torch.empty(max(k, 1), dtype=torch.float32), torch.empty(max(k, 1), dtype=torch.int64),
                torch.zeros(1, dtype=torch.int32))

    def local_search(self, d_query, k, record, lay, use_mask):
        ids, d, _, n_nan = ref.knn(self.rows, d_query.numpy(), k, rowids=self.rowids)
        buf = record.numpy()
        buf[lay.off_nan:lay.off_nan + 8].view(np.int64)[0] = n_nan
        buf[lay.off_rowids:lay.off_rowids + 8 * len(ids)].view(np.int64)[:] = ids
        buf[lay.off_dist:lay.off_dist + 4 * len(d)].view(np.float32)[:] = d
        buf[lay.off_count:lay.off_count + 4].view(np.int32)[0] = len(ids)

    def queries_to_device(self, queries):
        return torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32))

    def new_batch_outputs(self, nq, k):
        return (torch.empty((nq, max(k, 1)), dtype=torch.float32), torch.empty((nq, max(k, 1)), dtype=torch.int64),
                torch.zeros(nq, dtype=torch.int32))

    def local_search_batch(self, d_queries, k, record, lay):
        buf = record.numpy()
        for q, vec in enumerate(d_queries.numpy()):
            ids, d, _, n_nan = ref.knn(self.rows, vec, k, rowids=self.rowids)
            buf[lay.off_nan + 8 * q:lay.off_nan + 8 * q + 8].view(np.int64)[0] = n_nan
            o = lay.off_rowids + 8 * k * q
            buf[o:o + 8 * len(ids)].view(np.int64)[:] = ids
            o = lay.off_dist + 4 * k * q
            buf[o:o + 4 * len(d)].view(np.float32)[:] = d
            buf[lay.off_count + 4 * q:lay.off_count + 4 * q + 4].view(np.int32)[0] = len(ids)

    def merge_batch(self, gathered, nq, k, lay, out_dist, out_rowids, out_n):
        g = gathered.numpy()
        for q in range(nq):
            entries = []
            for l in range(g.shape[0]):
                cnt = int(g[l, lay.off_count + 4 * q:lay.off_count + 4 * q + 4].view(np.int32)[0])
                ids = g[l, lay.off_rowids + 8 * k * q:lay.off_rowids + 8 * k * (q + 1)].view(np.int64)
                d = g[l, lay.off_dist + 4 * k * q:lay.off_dist + 4 * k * (q + 1)].view(np.float32)
                entries += [(float(d[p]), l, p, int(ids[p])) for p in range(cnt)]
            entries.sort(key=lambda e: e[:3])
            entries = entries[:k]
            out_n[q] = len(entries)
            for i, e in enumerate(entries):
                out_dist[q, i] = e[0]
                out_rowids[q, i] = e[3]

    def merge(self, gathered, k, lay, out_dist, out_rowids, out_n):
        g = gathered.numpy()
        entries = []
        for l in range(g.shape[0]):
            cnt = int(g[l, lay.off_count:lay.off_count + 4].view(np.int32)[0])
            ids = g[l, lay.off_rowids:lay.off_rowids + 8 * k].view(np.int64)
            d = g[l, lay.off_dist:lay.off_dist + 4 * k].view(np.float32)
            entries += [(float(d[p]), l, p, int(ids[p])) for p in range(cnt)]
        entries.sort(key=lambda e: e[:3])
        entries = entries[:k]
        out_n[0] = len(entries)
        for i, e in enumerate(entries):
            out_dist[i] = e[0]
            out_rowids[i] = e[3]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = synth.unit_rows(n, DIM, 1234)
    rows[n - 1] = rows[3]                      # an exact tie across the first and last shard
    lo, hi = shard_bounds(n, world)[rank]
    index = ShardedIndex(OracleShardBackend(rows[lo:hi], lo + 1))
    queries = synth.unit_rows(3, DIM, 99)
    queries[2] = rows[3]
    got = [index.search(q, k) for q in queries]
    b_ids, b_d, b_n = index.search_batch(queries, k)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=np.stack([g[0] for g in got]),
             d=np.stack([g[1] for g in got]), b_ids=b_ids, b_d=b_d, b_n=b_n)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded_over_gloo(tmp_path, world):
    n, k = 5000, 20
    mp.spawn(_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    rows = synth.unit_rows(n, DIM, 1234)
    rows[n - 1] = rows[3]
    queries = synth.unit_rows(3, DIM, 99)
    queries[2] = rows[3]
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        for qi, q in enumerate(queries):
            ids, d, _, _ = ref.knn(rows, q, k, rowids=np.arange(1, n + 1))
            assert np.array_equal(got["ids"][qi], ids)
            assert np.array_equal(got["d"][qi], d)
            assert got["b_n"][qi] == k
            assert np.array_equal(got["b_ids"][qi], ids)        # batched sharded search == unsharded
            assert np.array_equal(got["b_d"][qi], d)
    # the planted tie comes back in rowid order across shards
    assert got["ids"][2][:2].tolist() == [4, n]


def test_record_layout_alignment():
    for k in (1, 2, 3, 20, 33, 100, 128):
        lay = RecordLayout(k)
        assert lay.off_rowids % 8 == 0 and lay.off_dist % 4 == 0 and lay.off_count % 4 == 0
        assert lay.nbytes % 16 == 0 and lay.nbytes >= lay.off_count + 4


def test_shard_bounds_cover_everything():
    for n, w in ((10, 3), (100_000_000, 8), (7, 8), (0, 2)):
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
