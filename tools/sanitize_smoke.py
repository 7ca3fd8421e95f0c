#!/usr/bin/env python3
"""Every kernel once at small sizes (small grids, 66k rows): a quick end-to-end exercise of the
single-query, batched, sign-code and fused-exchange paths, meant to be run under
`compute-sanitizer --tool memcheck|racecheck|synccheck` where that is allowed (it is closed on
the round-1 GPU pool), and as a plain smoke run otherwise:  python tools/sanitize_smoke.py"""
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_database_b200 import GpuIndex, synth  # noqa: E402


def main():
    import torch
    n = 66_000
    rows = synth.unit_rows(n, 1152, 1)
    rows[100] = 0
    q = synth.unit_rows(70, 1152, 2)
    with GpuIndex(0) as idx:
        idx.set_option("scan_ctas", 8)
        idx.load(rows)
        for k in (5, 100, 300):
            r = idx.search(q[0], k)
            assert r.counts[0] == k
        idx.set_option("scan_variant", 2)
        idx.search(q[1], 20)
        idx.set_option("scan_variant", 0)
        idx.search(q[1], 20, metric="l2")
        mask = np.arange(n) % 3 == 0
        idx.set_mask(mask)
        idx.search(q[2], 20, use_mask=True)
        idx.blend_search(q[3], 10, e2=q[4], weights=(0.7, 0.3), negatives=[q[5]], negative_weights=[0.5])
        idx.enable_batch()
        a = idx.search(q[:3], 20)
        b = idx.search(q, 100)
        idx.search(q[:3], 20, use_mask=True)
        codes = (rows >= 0).astype(np.uint8)
        idx.load_codes(codes)
        idx.binary_search((q[0] >= 0).astype(np.uint8), 20)
        idx.binary_search((q[0] >= 0).astype(np.uint8), 200)
        print("single/batch/binary ok", a.counts[:3], b.counts[:3])
    # fused exchange, two ranks on this device
    shards, inboxes = [], []
    for r in range(2):
        s = GpuIndex(0)
        s.load(rows[r * 33_000:(r + 1) * 33_000], np.arange(r * 33_000, (r + 1) * 33_000))
        s.set_option("scan_ctas", 4)
        s.set_option("xchg_timeout_ms", 120000)
        _, p = s.exchange_init(2, r)
        shards.append(s)
        inboxes.append(p)
    for s in shards:
        s.exchange_connect_pointers(inboxes, [0, 0])
    dq = torch.from_numpy(q[:1]).cuda()
    outs = [(torch.empty(20, dtype=torch.int64, device="cuda"), torch.empty(20, dtype=torch.float32, device="cuda"),
             torch.zeros(1, dtype=torch.int32, device="cuda")) for _ in range(2)]
    torch.cuda.synchronize()

    def one(r):
        shards[r].search_sharded_device(dq[0], 20, outs[r][0], outs[r][1], outs[r][2])
        shards[r].synchronize()
    ts = [threading.Thread(target=one, args=(r,)) for r in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    print("exchange ok", int(outs[0][2][0]), int(outs[1][2][0]))
    for s in shards:
        s.close()


if __name__ == "__main__":
    main()
