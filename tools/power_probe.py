#!/usr/bin/env python3
"""What does the board draw, and where does the SM clock settle, while (a) the float32 scan kernel, (b) a plain
read-only torch reduction over the same 46 GB, (c) the 64-query and (d) the 256-query contraction run back to back?
Samples NVML power / SM clock every 10 ms.  python tools/power_probe.py"""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml  # noqa: E402

from clip_database_b200 import GpuIndex  # noqa: E402


class Sampler:
    def __init__(self):
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.rows = []
        self._stop = threading.Event()

    def _run(self):
        while not self._stop.is_set():
            self.rows.append((pynvml.nvmlDeviceGetPowerUsage(self.h) / 1e3,
                              pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
            self._stop.wait(0.01)

    def measure(self, fn, seconds=2.0):
        self.rows = []
        self._stop.clear()
        t = threading.Thread(target=self._run, daemon=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t.start()
        n = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            n += 10
        dt = time.perf_counter() - t0
        self._stop.set()
        t.join()
        tail = self.rows[len(self.rows) // 2:]                      # the second half: steady state
        return {"iters": n, "ms_per_iter": dt * 1e3 / n, "power_w": float(np.median([r[0] for r in tail])),
                "sm_mhz": float(np.median([r[1] for r in tail]))}


def main():
    dev = torch.device("cuda", 0)
    n = 10_000_000
    rows = torch.empty((n, 1152), dtype=torch.float32, device=dev)
    for lo in range(0, n, 500_000):
        v = rows[lo:lo + 500_000]
        v.normal_()
        v.div_(v.norm(dim=1, keepdim=True))
    idx = GpuIndex(0)
    idx.attach(rows, rowid_base=1)
    idx.use_torch_stream()
    q = torch.randn((256, 1152), device=dev)
    o = (torch.empty((256, 100), dtype=torch.int64, device=dev), torch.empty((256, 100), dtype=torch.float32, device=dev),
         torch.zeros(256, dtype=torch.int32, device=dev), torch.zeros(256, dtype=torch.int64, device=dev),
         torch.zeros(256, dtype=torch.int32, device=dev))
    s = Sampler()
    out = {"idle": s.measure(lambda: None, 0.5)}
    out["scan_fp32"] = s.measure(lambda: idx.search_device(q[:1], 20, o[0][:1, :20], o[1][:1, :20], o[2][:1], o[3][:1]))
    out["scan_fp32"]["TBps"] = n * 4608 / 1e9 / out["scan_fp32"]["ms_per_iter"]
    flat = rows.view(-1)
    out["torch_sum_46GB"] = s.measure(lambda: flat.sum())
    out["torch_sum_46GB"]["TBps"] = n * 4608 / 1e9 / out["torch_sum_46GB"]["ms_per_iter"]
    idx.enable_batch()
    out["batch_nq1"] = s.measure(lambda: idx.search_batch_device(q[:1], 20, o[0][:1, :20], o[1][:1, :20], o[2][:1], o[3][:1], o[4][:1]))
    out["batch_nq256"] = s.measure(lambda: idx.search_batch_device(q, 100, *o))
    a = torch.randn((8192, 8192), device=dev, dtype=torch.bfloat16)
    b = torch.randn((8192, 8192), device=dev, dtype=torch.bfloat16)
    out["cublas_bf16_8192"] = s.measure(lambda: a @ b)
    out["cublas_bf16_8192"]["TFLOPs"] = 2 * 8192 ** 3 / 1e9 / out["cublas_bf16_8192"]["ms_per_iter"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
