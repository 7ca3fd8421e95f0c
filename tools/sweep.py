#!/usr/bin/env python3
"""Time the scan kernel alone (CUDA events inside the library) across its tuning knobs.
Usage (GPU box): python tools/sweep.py [--rows 10000000] [--reps 20]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import DIM, ROW_BYTES, generate_rows  # noqa: E402
from clip_database_b200 import GpuIndex  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = generate_rows(torch, dev, args.rows, 1234)
    idx = GpuIndex(0)
    idx.attach(rows, rowid_base=1)
    idx.use_torch_stream()
    q = torch.from_numpy(np.random.default_rng(99).standard_normal((8, DIM), dtype=np.float32)).to(dev)
    q /= q.norm(dim=1, keepdim=True)
    sms = idx.get_option("sm_count")
    results = []

    def run(label, k=20, **opts):
        for name, v in opts.items():
            idx.set_option(name, v)
        o_ids = torch.empty((1, max(k, 1)), dtype=torch.int64, device=dev)
        o_d = torch.empty((1, max(k, 1)), dtype=torch.float32, device=dev)
        o_n = torch.zeros(1, dtype=torch.int32, device=dev)
        for i in range(3):
            idx.search_device(q[i % 8].view(1, -1), k, o_ids, o_d, o_n)
        torch.cuda.synchronize()
        idx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.reps):
            idx.search_device(q[i % 8].view(1, -1), k, o_ids, o_d, o_n)
        e1.record()
        torch.cuda.synchronize()
        ms, n = idx.profile_read()
        idx.profile(False)
        scan = ms / n
        step = e0.elapsed_time(e1) / args.reps
        gb = args.rows * ROW_BYTES / 1e9
        rec = dict(label=label, k=k, opts=opts, scan_ms=scan, step_ms=step, scan_GBps=gb / scan * 1e3,
                   step_GBps=gb / step * 1e3)
        results.append(rec)
        print(json.dumps(rec), flush=True)

    base = dict(scan_variant=1, scan_ctas=0, evict_first=1, ldg_ctas_per_sm=4, scan_cfg=0, scan_assign=0,
                scan_chunk=4)
    run("tma default", **base)
    run("tma default (repeat)", **base)
    run("tma no-evict-hint", **{**base, "evict_first": 0})
    for cfg in (0, 1, 2, 3, 4):
        for assign in (0, 1, 2):
            run(f"tma cfg={cfg} assign={assign}", **{**base, "scan_cfg": cfg, "scan_assign": assign})
    for chunk in (1, 2, 8, 16):
        run(f"tma dynamic chunk={chunk}", **{**base, "scan_assign": 2, "scan_chunk": chunk})
    run("tma k=100", k=100, **base)
    run("tma k=100 dynamic", k=100, **{**base, "scan_assign": 2})
    run("general-k (k=1000)", k=1000, **base)
    run("ldg 6 ctas/sm", **{**base, "scan_variant": 2, "ldg_ctas_per_sm": 6})
    run("ldg 5 ctas/sm", **{**base, "scan_variant": 2, "ldg_ctas_per_sm": 5})
    run("tma default (end)", **base)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
