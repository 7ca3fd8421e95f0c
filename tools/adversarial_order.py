"""Worst case for the fused top-k: rows stored in order of decreasing distance to the query, so every
row is admitted by its warp's candidate list.  Prints the scan time against a random order (10M rows)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from bench import generate_rows, DIM
from clip_database_b200 import GpuIndex
dev = torch.device("cuda", 0)
rows = generate_rows(torch, dev, 10_000_000, 1234)
q = torch.from_numpy(np.random.default_rng(9).standard_normal(DIM, dtype=np.float32)).to(dev); q /= q.norm()
d = 1 - rows @ q
def timeit(idx, qq, k):
    idx.use_torch_stream()
    o = (torch.empty((1,k), dtype=torch.int64, device=dev), torch.empty((1,k), dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
    for _ in range(3): idx.search_device(qq.view(1,-1), k, *o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): idx.search_device(qq.view(1,-1), k, *o)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20
idx = GpuIndex(0); idx.attach(rows, rowid_base=0)
print("random order   k=20 %.3f ms  k=100 %.3f ms" % (timeit(idx, q, 20), timeit(idx, q, 100)))
idx.close()
order = torch.argsort(d, descending=True)
srt = rows[order]; del rows, order; torch.cuda.empty_cache()
idx = GpuIndex(0); idx.attach(srt, rowid_base=0)
print("farthest first k=20 %.3f ms  k=100 %.3f ms" % (timeit(idx, q, 20), timeit(idx, q, 100)))
