#!/usr/bin/env python3
"""Differential fuzz of the batched (tensor-core pre-selection + exact re-rank) path against the exact scan: random
store sizes, batch sizes, k, data shapes (unit rows, clusters, exact duplicates, scaled rows, zero rows), resident and
tiered stores, with and without a mask.  Every answer must be bit-identical.  python tools/fuzz_batch.py [--iters 150]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_database_b200 import GpuIndex  # noqa: E402

DIM = 1152


def make_rows(rng, n, kind):
    x = rng.standard_normal((n, DIM), dtype=np.float32)
    if kind == "clustered":
        centres = rng.standard_normal((max(1, n // 500), DIM), dtype=np.float32)
        x = centres[rng.integers(0, len(centres), n)] + 0.3 * x
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if kind == "scaled":
        x *= rng.uniform(0.01, 100.0, (n, 1)).astype(np.float32)
    if kind == "duplicates" and n > 10:
        src = rng.integers(0, n, n // 3)
        x[rng.integers(0, n, n // 3)] = x[src]
    if kind == "zeros" and n > 3:
        x[rng.integers(0, n, max(1, n // 50))] = 0
    return np.ascontiguousarray(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=150)
    ap.add_argument("--seed", type=int, default=2024)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    kinds = ["unit", "clustered", "scaled", "duplicates", "zeros"]
    bad = 0
    for it in range(args.iters):
        n = int(rng.choice([1, 7, 127, 128, 129, 255, 257, 1000, 5000, 33_000, 120_000]))
        n = max(1, n + int(rng.integers(-3, 4)))
        kind = kinds[it % len(kinds)]
        rows = make_rows(rng, n, kind)
        nq = int(rng.choice([1, 2, 3, 63, 64, 65, 128, 129, 200, 256]))
        k = int(min(n, rng.choice([1, 2, 5, 20, 32, 33, 64, 100, 128])))
        queries = make_rows(rng, nq, "unit")
        if it % 3 == 0:
            queries[0] = rows[int(rng.integers(0, n))]                      # a self match
        tiered = it % 4 == 1 and n >= 256
        use_mask = it % 5 == 2 and kind != "zeros"
        with GpuIndex(0) as idx:
            ids = np.arange(10, n + 10, dtype=np.int64)
            if tiered:
                idx.reserve(n, DIM, explicit_rowids=True, placement="host", device_rows=int(rng.integers(0, n)))
                for lo in range(0, n, 4000):
                    idx.append(rows[lo:lo + 4000], ids[lo:lo + 4000])
            else:
                idx.load(rows, ids)
            if use_mask:
                idx.set_mask(rng.random(n) < 0.7)
            idx.set_option("batch_min_nq", 1 << 20)
            want = idx.search(queries, k, use_mask=use_mask)
            idx.enable_batch()
            idx.set_option("batch_min_nq", 1)
            got = idx.search(queries, k, use_mask=use_mask)
        ok = (np.array_equal(got.counts, want.counts) and np.array_equal(got.rowids, want.rowids) and
              np.array_equal(got.distances.view(np.uint32), want.distances.view(np.uint32)) and
              np.array_equal(got.nan_rows, want.nan_rows))
        if not ok:
            bad += 1
            print("MISMATCH", dict(it=it, n=n, kind=kind, nq=nq, k=k, tiered=tiered, mask=use_mask), flush=True)
    print("fuzz: %d iterations, %d mismatches" % (args.iters, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
