#!/usr/bin/env python3
"""A/B a context option on the batched contraction inside ONE process, alternating the two settings
so that the board's power / thermal state (which moves pass B by +-15 % between runs) hits both.

    python tools/ab_option.py batch_static_tiles 1 0 [--batch 256] [--k 100] [--reps 6]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import DIM, generate_rows  # noqa: E402
from clip_database_b200 import GpuIndex  # noqa: E402


def ab_single(args, idx, dev):
    k = args.k
    q = torch.from_numpy(np.random.default_rng(9).standard_normal((8, DIM), dtype=np.float32)).to(dev)
    q /= q.norm(dim=1, keepdim=True)
    o = (torch.empty((1, k), dtype=torch.int64, device=dev), torch.empty((1, k), dtype=torch.float32, device=dev),
         torch.zeros(1, dtype=torch.int32, device=dev))
    res = {args.a: [], args.b: []}
    for _ in range(args.reps):
        for v in (args.a, args.b):
            idx.set_option(args.option, v)
            for i in range(5):
                idx.search_device(q[i % 8].view(1, -1), k, *o)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(60):
                idx.search_device(q[i % 8].view(1, -1), k, *o)
            e1.record()
            torch.cuda.synchronize()
            res[v].append(e0.elapsed_time(e1) / 60)
    for v in (args.a, args.b):
        print("%s=%d  ms per query %s  mean %.4f" % (args.option, v, [round(x, 3) for x in res[v]], float(np.mean(res[v]))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("option")
    ap.add_argument("a", type=int)
    ap.add_argument("b", type=int)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--single", action="store_true", help="time the single-query float32 scan instead of a batch")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = generate_rows(torch, dev, args.rows, 1234)
    idx = GpuIndex(0)
    idx.attach(rows, rowid_base=1)
    idx.use_torch_stream()
    if args.single:
        return ab_single(args, idx, dev)
    idx.enable_batch()
    B, k = args.batch, args.k
    q = torch.from_numpy(np.random.default_rng(9).standard_normal((B, DIM), dtype=np.float32)).to(dev)
    q /= q.norm(dim=1, keepdim=True)
    o = (torch.empty((B, k), dtype=torch.int64, device=dev), torch.empty((B, k), dtype=torch.float32, device=dev),
         torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int64, device=dev),
         torch.zeros(B, dtype=torch.int32, device=dev))
    res = {args.a: [], args.b: []}
    step = {args.a: [], args.b: []}
    for _ in range(args.reps):
        for v in (args.a, args.b):
            idx.set_option(args.option, v)
            for _ in range(3):
                idx.search_batch_device(q, k, *o)
            torch.cuda.synchronize()
            idx.profile(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                idx.search_batch_device(q, k, *o)
            e1.record()
            torch.cuda.synchronize()
            ms, n = idx.profile_read()
            idx.profile(False)
            res[v].append(ms / n)
            step[v].append(e0.elapsed_time(e1) / 20)
    for v in (args.a, args.b):
        print("%s=%d  pass B ms %s  mean %.3f | step ms mean %.3f" % (
            args.option, v, [round(x, 3) for x in res[v]], float(np.mean(res[v])), float(np.mean(step[v]))))


if __name__ == "__main__":
    main()
