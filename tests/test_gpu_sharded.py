"""Row-sharded search over NCCL on real GPUs (needs >= 2 devices; skipped on a 1-GPU box)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from clip_database_b200 import synth
from oracle import ref

from conftest import have_gpu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from clip_database_b200 import GpuIndex, synth
from clip_database_b200.sharded import CudaShardBackend, ShardedIndex, shard_bounds
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n, k = 40_000, 20
rows = synth.unit_rows(n, 1152, 1234)
rows[n - 1] = rows[3]
lo, hi = shard_bounds(n, world)[rank]
idx = GpuIndex(rank)
idx.load(rows[lo:hi], np.arange(lo + 1, hi + 1))
backend = CudaShardBackend(idx)
sh = ShardedIndex(backend, fused=True)         # scan + peer-memory exchange + merge in one launch per rank
assert sh.fused
sh_nccl = ShardedIndex(CudaShardBackend(idx), fused=False)   # local search + NCCL all-gather + merge kernel
queries = synth.unit_rows(4, 1152, 99)
queries[3] = rows[3]
out = []
for q in queries:
    before = idx.launch_count
    ids, d = sh.search(q, k)
    assert idx.launch_count - before == 1, "fused sharded search must be ONE launch"
    nan = sh.nan_rows()
    ids2, d2 = sh_nccl.search(q, k)
    assert np.array_equal(ids, ids2) and np.array_equal(d.view(np.uint32), d2.view(np.uint32))
    assert nan == sh_nccl.nan_rows()
    out.append((ids.tolist(), d.tolist(), nan))
m_ids, m_d, m_n = sh.search_batch(queries, k)          # no bf16 store on this index: one fused scan per query
assert np.array_equal(m_ids[0], out[0][0]) and np.array_equal(m_ids[3], out[3][0]) and m_n.tolist() == [k] * 4
ids100, d100 = sh.search(queries[0], 100)
ids100n, d100n = sh_nccl.search(queries[0], 100)
assert np.array_equal(ids100, ids100n) and np.array_equal(d100.view(np.uint32), d100n.view(np.uint32))
n2 = 140_000
rows2 = synth.unit_rows(n2, 1152, 77)
lo2, hi2 = shard_bounds(n2, world)[rank]
idx2 = GpuIndex(rank)
idx2.load(rows2[lo2:hi2], np.arange(lo2 + 1, hi2 + 1))
idx2.enable_batch()
sh2 = ShardedIndex(CudaShardBackend(idx2), fused=True)     # candidates exchanged by the batched path's last kernel
sh2_nccl = ShardedIndex(CudaShardBackend(idx2), fused=False)  # one NCCL all-gather + merge kernel
bq = synth.unit_rows(40, 1152, 5)
bq[7] = 0                                       # flagged by every shard -> re-run through the exact sharded search
b_ids, b_d, b_n = sh2.search_batch(bq, 50)
n_ids, n_d, n_n = sh2_nccl.search_batch(bq, 50)
assert np.array_equal(b_ids, n_ids) and np.array_equal(b_n, n_n)
assert np.array_equal(b_d.view(np.uint32), n_d.view(np.uint32))
big = synth.unit_rows(300, 1152, 6)             # two passes (256 + 44)
g_ids, g_d, g_n = sh2.search_batch(big, 20)
h_ids, h_d, h_n = sh2_nccl.search_batch(big, 20)
assert np.array_equal(g_ids, h_ids) and np.array_equal(g_d.view(np.uint32), h_d.view(np.uint32))
np.savez(os.path.join({out!r}, f"batch{{rank}}.npz"), ids=b_ids, d=b_d, n=b_n)
idx2.close()
json.dump(out, open(os.path.join({out!r}, f"rank{{rank}}.json"), "w"))
idx.close()
dist.destroy_process_group()
"""


def test_nccl_sharded_search_equals_unsharded(tmp_path):
    assert have_gpu()
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=str(tmp_path)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-3000:]
    n, k = 40_000, 20
    rows = synth.unit_rows(n, 1152, 1234)
    rows[n - 1] = rows[3]
    queries = synth.unit_rows(4, 1152, 99)
    queries[3] = rows[3]
    from conftest import assert_topk_parity
    for rank in range(world):
        got = json.load(open(tmp_path / f"rank{rank}.json"))
        for qi, q in enumerate(queries):
            ids, d, nan = got[qi]
            _, od, oseq, _ = ref.knn(rows, q, k)
            assert nan == 0
            assert_topk_parity(np.array(ids) - 1, np.array(d, dtype=np.float32), oseq, od, ref.distances(rows, q))
        assert got[3][0][:2] == [4, n]          # cross-shard exact tie in rowid order
    # batched sharded search (tensor-core path per shard) == single-store exact search
    rows2 = synth.unit_rows(140_000, 1152, 77)
    bq = synth.unit_rows(40, 1152, 5)
    bq[7] = 0
    from clip_database_b200 import GpuIndex
    with GpuIndex(0) as whole:
        whole.load(rows2, np.arange(1, rows2.shape[0] + 1))
        want = whole.search(bq, 50)
    for rank in range(world):
        b = np.load(tmp_path / f"batch{rank}.npz")
        assert np.array_equal(b["n"], want.counts)
        assert np.array_equal(b["ids"], want.rowids)
        assert np.array_equal(b["d"].view(np.uint32), want.distances.view(np.uint32))
    # every rank holds the same merged answer
    assert json.load(open(tmp_path / "rank0.json")) == json.load(open(tmp_path / f"rank{world - 1}.json"))
