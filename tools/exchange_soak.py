#!/usr/bin/env python3
"""Soak test of the fused shard exchange: thousands of back-to-back sharded searches with changing k
(so both inbox banks, all list lengths and the epoch wrap of the bank parity are exercised), every
answer compared with the NCCL all-gather path.  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node 2 tools/exchange_soak.py [--iters 3000]"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_database_b200 import GpuIndex, synth  # noqa: E402
from clip_database_b200.sharded import CudaShardBackend, ShardedIndex, shard_bounds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3000)
    ap.add_argument("--rows", type=int, default=120_000)
    ap.add_argument("--wrap", action="store_true",
                    help="start both sequence counters just below 2^32 so the run crosses the uint32 wrap")
    ap.add_argument("--batch-every", type=int, default=7, help="every n-th iteration is a 48-query batched search")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    rows = synth.unit_rows(args.rows, 1152, 1234)
    lo, hi = shard_bounds(args.rows, world)[rank]
    idx = GpuIndex(rank)
    idx.load(rows[lo:hi], np.arange(lo + 1, hi + 1))
    idx.enable_batch()
    fused = ShardedIndex(CudaShardBackend(idx), fused=True)
    nccl = ShardedIndex(CudaShardBackend(idx), fused=False)
    if args.wrap:
        idx.exchange_set_epoch(0xFFFFFF00, 0xFFFFFFF0)      # single-query counter wraps after 255 searches, batched after 15
    queries = synth.unit_rows(64, 1152, 99)
    rng = np.random.default_rng(7)                       # same sequence on every rank
    ks = rng.choice([1, 5, 20, 33, 100, 128], size=args.iters)
    bad = 0
    for it in range(args.iters):
        q, k = queries[it % 64], int(ks[it])
        if args.batch_every and it % args.batch_every == 0:
            qs = queries[(it % 16):(it % 16) + 48]
            a = fused.search_batch(qs, min(k, 100))
            if it % (10 * args.batch_every) == 0:
                b = nccl.search_batch(qs, min(k, 100))
                if not (np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))):
                    bad += 1
            continue
        a_ids, a_d = fused.search(q, k)
        if it % 10 == 0:                                 # the reference answer every 10th iteration
            b_ids, b_d = nccl.search(q, k)
            if not (np.array_equal(a_ids, b_ids) and np.array_equal(a_d.view(np.uint32), b_d.view(np.uint32))):
                bad += 1
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("exchange soak: %d searches on %d ranks, %d mismatches" % (args.iters, world, int(t[0])))
    idx.close()
    dist.destroy_process_group()
    sys.exit(1 if int(t[0]) else 0)


if __name__ == "__main__":
    main()
