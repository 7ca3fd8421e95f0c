#!/usr/bin/env python3
"""bench.py — KNN scan throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

A "step" is one pass of the hot path over one query: scan every resident
1152-d float32 row with cosine distance and select the top k (k = 20).
  N = 1   BASELINE configs[1]: 10M rows (46.08 GB) resident on one B200.
  N > 1   row-sharded (weak scaling): 12.5M rows per GPU, i.e. the 100M-row
          configs[4] at N = 8; per-rank scan + one NCCL all-gather of k
          candidates per rank + merge on every rank.
`value` is whole-job scan GB/s with the query already in HBM (algorithmic bytes
= rows * 1152 * 4 per query); `e2e` is the same metric through
the public host API (host query in, host results out, copies inside the timed
region).  Inputs (46 GB per scan) are far larger than L2 (126 MB), so no flush
is needed between steps.

The reference arm times the reference's statement (image_database.py:1564-1574)
on the real SQLite with the oracle's C restatement of sqlite-vec's
vec_distance_cosine, single-threaded like the reference, on a bounded row
sample; the same thing is reported as `cpu_baseline` by the default arm.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIM = 1152
ROW_BYTES = DIM * 4
METRIC = "knn_scan_throughput"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default 10M at N=1, 12.5M at N>1)")
    ap.add_argument("--workload", default="single", choices=["single", "blend", "batch", "binary"],
                    help="single: configs[1]; blend: configs[3] (0.7/0.3 blend + negative, then the scan); "
                         "batch: configs[2] (B queries per step through the tcgen05 contraction + fp32 re-rank); "
                         "binary: the sign-code fallback search (SURVEY §8 f-4) over bit-packed codes")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--data", default="uniform", choices=["uniform", "clustered"],
                    help="uniform: seeded random unit rows and queries (BASELINE configs); clustered: 1024 clusters, "
                         "queries near stored rows (SURVEY §8d variant)")
    ap.add_argument("--sample-stride", type=int, default=0, help="batch: pass A samples 1/s of the rows (0 = auto)")
    ap.add_argument("--no-refine", action="store_true", help="batch: skip the second threshold")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fused", "nccl"],
                    help="N > 1: fused = the scan kernel's last CTA exchanges candidates over NVLink peer memory "
                         "and merges (one launch per rank); nccl = all-gather + merge kernel; auto = fused when "
                         "every rank can map its peers' memory, else nccl")
    ap.add_argument("--variant", type=int, default=0, help="scan kernel: 0 auto, 1 TMA ring, 2 direct loads")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=100_000)
    ap.add_argument("--cpu-sample-queries", type=int, default=40)
    return ap.parse_args()


def workload_config(args, rows_per_gpu, n_gpus):
    cfg = {
        "workload": ("single-query cosine KNN, k=%d, %d x 1152 fp32 rows resident per GPU "
                     "(BASELINE configs[1])" % (args.k, rows_per_gpu)) if n_gpus == 1 else
                    ("row-sharded single-query cosine KNN, k=%d, %d x 1152 fp32 rows per GPU x %d GPUs "
                     "(BASELINE configs[4] shard size; 100M rows at 8 GPUs)" % (args.k, rows_per_gpu, n_gpus)),
        "rows_per_gpu": rows_per_gpu,
        "rows_total": rows_per_gpu * n_gpus,
        "dim": DIM,
        "k": args.k,
        "metric_kind": "cosine",
        "blend": args.workload == "blend",
        "l2": "inputs_larger_than_L2",
        "parallelism": "row-shard x%d" % n_gpus if n_gpus > 1 else "single GPU",
    }
    if getattr(args, "exchange_used", None):
        cfg["exchange"] = args.exchange_used
    if n_gpus in (2, 4):
        cfg["note"] = "100M fp32 rows do not fit %d GPUs (%.1f GB/GPU); weak-scaled shard of 12.5M rows/GPU" % (
            n_gpus, 100e6 * ROW_BYTES / n_gpus / 1e9)
    return cfg


# ---------------------------------------------------------------- clocks ------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""
    REASONS = {
        0x0000000000000004: "sw_power_cap", 0x0000000000000008: "hw_slowdown",
        0x0000000000000020: "sw_thermal_slowdown", 0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000080: "hw_power_brake_slowdown", 0x0000000000000002: "applications_clocks_setting",
        0x0000000000000100: "display_clock_setting", 0x0000000000000010: "sync_boost",
    }

    def __init__(self, device_index: int):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    u = pynvml.nvmlDeviceGetUUID(h)
                    u = u.decode() if isinstance(u, bytes) else u
                    if uuid in u:
                        self.h = h
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:          # no NVML: report that instead of inventing clocks
            self.nv = None
            self.error = repr(e)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) \
                    if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.error}
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------- CPU arm -----------------
def cpu_reference_path(sample_rows: int, n_queries: int, warmup: int = 2):
    """The reference's CPU path on a bounded sample: its SQL statement executed by the real
    SQLite (1 thread, like the reference) with vec_distance_cosine from oracle/vec_shim.so.
    Times the interval the reference itself calls `db_query` (image_database.py:1557-1631).
    Returns (seconds per query, provider string, extra dict)."""
    from clip_database_b200 import synth
    from oracle import ref, sql_harness

    rows = ref.fill_unit_rows(sample_rows, DIM, 1234)
    queries = ref.fill_unit_rows(max(n_queries, 1) + warmup, DIM, 99)
    tmp = tempfile.mkdtemp(prefix="clipdb_cpu_")
    db_path = os.path.join(tmp, "sample.db")
    synth.write_reference_db(db_path, rows)
    conn, provider = sql_harness.connect(db_path)
    for q in queries[:warmup]:
        sql_harness.run_statement(conn, q, 20)
    times = []
    for q in queries[warmup:]:
        t0 = time.perf_counter()
        sql_harness.run_statement(conn, q, 20)
        times.append(time.perf_counter() - t0)
    conn.close()
    # the bare scalar loop + bounded top-k (no SQLite machinery): an upper bound on what the
    # reference's single thread could reach, and the same loop on every host core
    t0 = time.perf_counter()
    for q in queries[warmup:warmup + 8]:
        ref.knn(rows, q, 20)
    loop_s = (time.perf_counter() - t0) / min(8, len(queries) - warmup)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    for q in queries[warmup:warmup + 8]:
        ref.distances(rows, q, threads=cores)
    mt_s = (time.perf_counter() - t0) / min(8, len(queries) - warmup)
    try:
        os.remove(db_path)
        os.rmdir(tmp)
    except OSError:
        pass
    gb = sample_rows * ROW_BYTES / 1e9
    extra = {"scalar_loop_1thread_GBps": gb / loop_s, "scalar_loop_all_cores_GBps": gb / mt_s,
             "host_cores": cores}
    return float(np.mean(times)), provider, extra


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_gpus = args.gpus
    rows_per_gpu = args.rows or (10_000_000 if n_gpus == 1 else 12_500_000)
    # bound the sample so warmup + steps statements end within about a minute
    per_row = 2.0e-6
    sample = int(min(args.cpu_sample_rows, max(10_000, 60.0 / max(args.steps + args.warmup, 1) / per_row)))
    sec, provider, extra = cpu_reference_path(sample, args.steps, warmup=max(args.warmup, 1))
    gbps = sample * ROW_BYTES / 1e9 / sec
    sample_desc = ("reference SQL statement (image_database.py:1564-1574) on SQLite via %s; %d-row sample "
                   "of the workload, %d timed queries, k=20, 1 thread (SQLite runs a statement on one "
                   "thread); vec0 is a plain stand-in table" % (provider, sample, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": gbps, "unit": UNIT, "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, rows_per_gpu, n_gpus),
        "queries_per_s_at_sample": 1.0 / sec,
        "queries_per_s_extrapolated_to_config": 1.0 / (sec * rows_per_gpu * n_gpus / sample),
        "cpu_baseline": {"value": gbps, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample_desc,
                         "cpu": cpu_model(), **extra},
        "e2e": {"value": gbps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------- GPU arm -----------------
N_CLUSTERS = 1024


def generate_rows(torch, device, n_rows, seed, clustered=False):
    """randn float32 from a seeded CUDA generator, in chunks,
    rows L2-normalised, written straight into the resident matrix.  ``clustered``: the §8d
    variant — row r = normalise(centre[r % 1024] + noise) with |noise| ~ 0.75 |centre| (cosine
    ~0.8 inside a cluster), so the top-k of a query near a stored row is a dense neighbourhood
    instead of the tail of a noise distribution."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    rows = torch.empty((n_rows, DIM), dtype=torch.float32, device=device)
    centres = None
    if clustered:
        cgen = torch.Generator(device=device)
        cgen.manual_seed(4242)                      # the same centres on every rank
        centres = torch.randn((N_CLUSTERS, DIM), generator=cgen, device=device)
        centres /= centres.norm(dim=1, keepdim=True)
    chunk = 500_000
    for lo in range(0, n_rows, chunk):
        hi = min(lo + chunk, n_rows)
        view = rows[lo:hi]
        view.normal_(generator=gen)
        if clustered:
            view.mul_(0.75 / DIM ** 0.5)
            view.add_(centres[torch.arange(lo, hi, device=device) % N_CLUSTERS])
        view.div_(view.norm(dim=1, keepdim=True))
    return rows


def make_queries(rows_t, n_q, seed, clustered):
    """Host float32 unit queries: seeded noise, or (clustered) stored rows + 10 % noise."""
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((n_q, DIM), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    if clustered:
        picks = rng.choice(rows_t.shape[0], n_q, replace=False)
        base = rows_t[picks.tolist()].cpu().numpy()
        q = base + 0.1 * q
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.ascontiguousarray(q, dtype=np.float32)


def run_batch_workload(args, torch, idx, rows, rows_per_gpu, device, rank, local_rank):
    """BASELINE configs[2]: B = 256 queries per step, k = 100 (pass --k 100), one GPU.
    metric = queries/s; roofline = tensor pipe (flops of the full-store contraction / the
    filter-pass kernel's own duration)."""
    B, k = args.batch, args.k
    idx.enable_batch()
    idx.set_option("batch_sample_stride", args.sample_stride)
    idx.set_option("batch_refine", 0 if args.no_refine else 1)
    idx.set_option("batch_min_nq", 1)          # --batch 1: one query through the bf16 pre-selection
    n_sets = 4
    host_q = make_queries(rows, n_sets * B, 99, args.data == "clustered").reshape(n_sets, B, DIM)
    d_q = torch.from_numpy(host_q).to(device)
    o_ids = torch.empty((B, k), dtype=torch.int64, device=device)
    o_dist = torch.empty((B, k), dtype=torch.float32, device=device)
    o_n = torch.zeros(B, dtype=torch.int32, device=device)
    o_nan = torch.zeros(B, dtype=torch.int64, device=device)
    flags = torch.zeros(B, dtype=torch.int32, device=device)
    sampler = ClockSampler(local_rank)

    for i in range(args.warmup):
        idx.search_batch_device(d_q[i % n_sets], k, o_ids, o_dist, o_n, o_nan, flags)
    torch.cuda.synchronize()
    sampler.start()
    launches0 = idx.launch_count
    idx.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        idx.search_batch_device(d_q[i % n_sets], k, o_ids, o_dist, o_n, o_nan, flags)
    ev1.record()
    torch.cuda.synchronize()
    ms_step = ev0.elapsed_time(ev1) / args.steps
    gemm_ms, gemm_n = idx.profile_read()
    idx.profile(False)
    launches = idx.launch_count - launches0
    flagged = int((flags != 0).sum())
    cand, surv = idx.batch_stats()

    for i in range(min(args.warmup, 3)):
        idx.search(host_q[i % n_sets], k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        last = idx.search(host_q[i % n_sets], k)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    sampler.stop()
    assert np.all(last.counts == k)

    # spot check against the exact single-query path (same context, batch path bypassed by nq=1)
    idx.set_option("batch_min_nq", 1 << 20)
    check = [idx.search(host_q[(args.steps - 1) % n_sets][j], k) for j in (0, B // 2, B - 1)]
    for j, c in zip((0, B // 2, B - 1), check):
        assert np.array_equal(c.rowids[0], last.rowids[j]) and np.array_equal(c.distances[0], last.distances[j])

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flops = 2.0 * B * rows_per_gpu * DIM
    # the same box in the same power state: cuBLAS bf16 GEMM back to back for ~1 s right after the
    # timed region (the board sits at its power cap whenever the tensor pipe is busy, and the clock
    # it settles at varies from box to box and minute to minute)
    live = None
    try:
        ga = torch.randn((8192, 8192), device=device, dtype=torch.bfloat16)
        gb = torch.randn((8192, 8192), device=device, dtype=torch.bfloat16)
        for _ in range(10):
            ga @ gb
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 600
        g0.record()
        for _ in range(reps):
            ga @ gb
        g1.record()
        torch.cuda.synchronize()
        live = 2.0 * 8192 ** 3 * reps / 1e12 / (g0.elapsed_time(g1) / 1e3)
        del ga, gb
    except Exception:
        pass
    gemm_avg = gemm_ms / max(gemm_n, 1)
    tflops = flops / 1e12 / (gemm_avg / 1e3)
    store_gbps = rows_per_gpu * DIM * 2 / 1e9 / (gemm_avg / 1e3)
    if B > 128:     # 256 queries per pass: AI = 256 flop/B, above the ridge -> tensor-bound
        roof = {"bound": "tensor", "kernel": "batch_gemm_pair_kernel<FILTER, 256>", "achieved": tflops, "peak": peak,
                "unit": "TFLOP/s", "frac": tflops / peak, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained",
                "flops_per_launch": flops, "avg_launch_ms": gemm_avg, "launches_timed": int(gemm_n),
                "hbm_GBps_bf16_store": store_gbps, "live_cublas_tflops": live,
                "frac_of_live_cublas": (tflops / live) if live else None,
                "traffic": 23040801000.0 + 9860352.0 if rows_per_gpu == 10_000_000 and B == 256 else None,
                "traffic_source": "profiles/r01v4_gemm_ncu_raw.csv (dram read + write, one launch)"}
    else:           # 64 / 128 queries per pass: the contraction streams the bf16 store -> HBM-bound
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        roof = {"bound": "hbm", "kernel": "batch_gemm_pair_kernel<FILTER, %d>" % (64 if B <= 64 else 128),
                "achieved": store_gbps, "peak": hbm_peak, "unit": "GB/s", "frac": store_gbps / hbm_peak,
                "frac_of_nominal_8000": store_gbps / 8000.0, "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)",
                "algorithmic_bytes_per_launch": rows_per_gpu * DIM * 2, "avg_launch_ms": gemm_avg,
                "launches_timed": int(gemm_n), "tensor_TFLOPs": tflops, "traffic": None}
    line = {
        "metric": "knn_batched_queries_per_s", "value": B * 1e3 / ms_step, "unit": "queries/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16 pre-select + f32 re-rank",
        "data": "synthetic" if args.data == "uniform" else "synthetic, clustered (1024 clusters, queries near stored rows)",
        "config": {"workload": "batched cosine KNN, B=%d queries per step, k=%d, %d x 1152 rows "
                               "(BASELINE configs[2]): tcgen05 contraction + fp32 re-rank" % (B, k, rows_per_gpu),
                   "rows_per_gpu": rows_per_gpu, "batch": B, "k": k, "dim": DIM,
                   "store_bytes": {"fp32": rows_per_gpu * ROW_BYTES, "bf16": rows_per_gpu * DIM * 2},
                   "l2": "inputs_larger_than_L2", "sample_stride": args.sample_stride,
                   "refine": not args.no_refine},
        "equivalent_scan_GBps_fp32": B * rows_per_gpu * ROW_BYTES / 1e9 / (ms_step / 1e3),
        "e2e": {"value": B * 1e3 / e2e_ms, "unit": "queries/s", "h2d_bytes_per_step": B * ROW_BYTES,
                "d2h_bytes_per_step": B * (k * 12 + 12) + B * 4, "ms_per_step": e2e_ms,
                "api": "GpuIndex.search (clipdb_search, nq=%d)" % B},
        "gpu_launches": int(launches), "flagged_queries_last_step": flagged,
        "candidates_per_query": {"filter_mean": float(cand[:B].mean()), "filter_max": int(cand[:B].max()),
                                 "reranked_mean": float(surv[:B].mean()), "reranked_max": int(surv[:B].max())},
        "roofline": roof,
        "clocks": sampler.summary(),
    }
    emit(line)
    idx.close()
    return 0


def run_binary_workload(args, torch, device, local_rank):
    """one query's AND-popcount scan + top-k over N bit-packed sign codes (144 B per
    row resident).  metric = scanned GB/s of the packed store; HBM roofline."""
    from clip_database_b200 import GpuIndex
    k = args.k
    n = args.rows or 10_000_000
    idx = GpuIndex(local_rank)
    idx.use_torch_stream()
    gen = torch.Generator(device=device)
    gen.manual_seed(1234)
    # codes are generated and packed chunk by chunk (n x 1152 bytes would be 11.5 GB at 10M rows)
    chunk = 1_000_000
    parts = []
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        parts.append((torch.randn((m, DIM), generator=gen, device=device) >= 0).to(torch.uint8))
    codes = torch.cat(parts)
    del parts
    idx.load_codes(codes)
    del codes
    torch.cuda.empty_cache()
    rng = np.random.default_rng(99)
    n_q = 64
    host_codes = (rng.standard_normal((n_q, DIM), dtype=np.float32) >= 0).astype(np.uint8)
    words = np.zeros((n_q, DIM // 32), dtype=np.uint32)
    words.view(np.uint8)[:] = np.packbits(host_codes, axis=1, bitorder="little")
    d_words = torch.from_numpy(words.view(np.int32)).to(device)
    o_ids = torch.empty(k, dtype=torch.int64, device=device)
    o_sc = torch.empty(k, dtype=torch.int32, device=device)
    o_n = torch.zeros(1, dtype=torch.int32, device=device)
    sampler = ClockSampler(local_rank)
    for i in range(args.warmup):
        idx.binary_search_device(d_words[i % n_q], k, o_ids, o_sc, o_n)
    torch.cuda.synchronize()
    sampler.start()
    launches0 = idx.launch_count
    idx.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        idx.binary_search_device(d_words[i % n_q], k, o_ids, o_sc, o_n)
    ev1.record()
    torch.cuda.synchronize()
    ms_step = ev0.elapsed_time(ev1) / args.steps
    scan_ms, scans = idx.profile_read()
    idx.profile(False)
    launches = idx.launch_count - launches0
    for i in range(3):
        idx.binary_search(host_codes[i], k)
    t0 = time.perf_counter()
    for i in range(args.steps):
        last = idx.binary_search(host_codes[i % n_q], k)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    sampler.stop()
    assert len(last[0]) == k and np.all(np.diff(last[1]) <= 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    gb = n * (DIM // 8) / 1e9
    scan_avg = scan_ms / max(scans, 1)
    line = {
        "metric": "binary_scan_throughput", "value": gb / (ms_step / 1e3), "unit": "GB/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 (AND + popcount)", "data": "synthetic",
        "config": {"workload": "sign-code fallback search (image_database.py:1591-1629), k=%d, %d x 1152-bit codes "
                               "(144 B per row resident), score = popcount(q AND row) mod 256" % (k, n),
                   "rows_per_gpu": n, "k": k, "dim": DIM, "l2": "inputs_larger_than_L2"},
        "queries_per_s": 1e3 / ms_step, "rows_per_s": n / (ms_step / 1e3),
        "e2e": {"value": gb / (e2e_ms / 1e3), "unit": "GB/s", "h2d_bytes_per_step": DIM // 8,
                "d2h_bytes_per_step": k * 12 + 12, "ms_per_step": e2e_ms, "queries_per_s": 1e3 / e2e_ms,
                "api": "GpuIndex.binary_search (clipdb_binary_search)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "binary_scan_kernel", "achieved": gb / (scan_avg / 1e3), "peak": peak,
                     "unit": "GB/s", "frac": gb / (scan_avg / 1e3) / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)",
                     "algorithmic_bytes_per_launch": n * (DIM // 8), "avg_launch_ms": scan_avg,
                     "launches_timed": int(scans), "traffic": None},
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline:
        # the reference's own fallback on the host: fetch every row, np.frombuffer + np.dot per row
        from clip_database_b200 import synth
        from oracle import ref, sql_harness
        m = min(args.cpu_sample_rows, 50_000)
        rows = ref.fill_unit_rows(m, DIM, 1234)
        tmp = tempfile.mkdtemp(prefix="clipdb_bin_")
        db_path = os.path.join(tmp, "b.db")
        synth.write_reference_db(db_path, rows, vectors=False)
        q = ref.fill_unit_rows(4, DIM, 99)
        sql_harness.reference_binary_search(db_path, q[0], k)
        t0 = time.perf_counter()
        for j in range(1, 4):
            sql_harness.reference_binary_search(db_path, q[j], k)
        sec = (time.perf_counter() - t0) / 3
        os.remove(db_path)
        os.rmdir(tmp)
        line["cpu_baseline"] = {"value": m * (DIM // 8) / 1e9 / sec, "unit": "GB/s", "cores": 1, "kind": "port",
                                "rows_per_s": m / sec,
                                "sample": "literal restatement of image_database.py:1591-1629 (real SQLite fetchall + "
                                          "np.frombuffer + np.dot per row, 1 thread), %d-row sample, 3 queries; GB/s in "
                                          "the same packed-store bytes (the reference itself reads 1152 B per row)" % m,
                                "cpu": cpu_model()}
    emit(line)
    idx.close()
    return 0


def run_sharded_batch_workload(args, torch, dist, idx, sharded, rows_per_gpu, device, rank, local_rank, world):
    """BASELINE configs[4] with configs[2]'s queries: B queries per step against a store row-sharded
    over `world` GPUs.  Per rank: tensor-core batched search of its shard; ONE NCCL all-gather of
    B*k candidates per rank; per-query merge on every rank.  metric = queries/s (whole job)."""
    B, k = args.batch, args.k
    idx.enable_batch()
    rng = np.random.default_rng(99)
    n_sets = 4
    host_q = rng.standard_normal((n_sets, B, DIM), dtype=np.float32)
    host_q /= np.linalg.norm(host_q, axis=2, keepdims=True)
    d_q = torch.from_numpy(host_q).to(device)
    sampler = ClockSampler(local_rank) if rank == 0 else None

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    for i in range(args.warmup):
        sharded.search_batch_device(d_q[i % n_sets], k)
    barrier()
    if sampler:
        sampler.start()
    launches0 = idx.launch_count
    idx.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        out = sharded.search_batch_device(d_q[i % n_sets], k)
    ev1.record()
    barrier()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    gemm_ms, gemm_n = idx.profile_read()
    idx.profile(False)
    launches = idx.launch_count - launches0
    flagged = int((out[3].contiguous().view(torch.int32) != 0).sum())
    for i in range(min(args.warmup, 3)):
        sharded.search_batch(host_q[i % n_sets], k)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        last = sharded.search_batch(host_q[i % n_sets], k)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    if sampler:
        sampler.stop()
    assert np.all(last[2] == k) and np.all(np.diff(last[1], axis=1) >= 0)
    # spot check against the exact single-query sharded search (fused exchange / NCCL)
    j = B // 3
    one_ids, one_d = sharded.search(host_q[(args.steps - 1) % n_sets][j], k)
    assert np.array_equal(one_ids, last[0][j]) and np.array_equal(one_d.view(np.uint32), last[1][j].view(np.uint32))
    gemm_avg = max_over_ranks(gemm_ms / max(gemm_n, 1))
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        flops = 2.0 * B * rows_per_gpu * DIM
        emit({
            "metric": "knn_batched_queries_per_s", "value": B * 1e3 / ms_step, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 pre-select + f32 re-rank", "data": "synthetic",
            "config": {"workload": "row-sharded batched cosine KNN, B=%d queries per step, k=%d, %d x 1152 rows per GPU "
                                   "x %d GPUs (BASELINE configs[4] store, configs[2] queries)" % (B, k, rows_per_gpu, world),
                       "rows_per_gpu": rows_per_gpu, "rows_total": rows_per_gpu * world, "batch": B, "k": k, "dim": DIM,
                       "l2": "inputs_larger_than_L2", "parallelism": "row-shard x%d" % world,
                       "exchange": ("peer-memory exchange in the batched path's last kernel" if sharded.fused else
                                    "one NCCL all-gather of B*k candidates per rank + merge kernel")},
            "e2e": {"value": B * 1e3 / e2e_ms, "unit": "queries/s", "h2d_bytes_per_step": B * ROW_BYTES,
                    "d2h_bytes_per_step": B * (k * 12 + 4) + 4 * B * world, "ms_per_step": e2e_ms,
                    "api": "ShardedIndex.search_batch"},
            "gpu_launches": int(launches), "flagged_queries_last_step": flagged,
            "roofline": ({"bound": "tensor", "kernel": "batch_gemm_pair_kernel<FILTER, 256>",
                          "achieved": flops / 1e12 / (gemm_avg / 1e3), "peak": peak, "unit": "TFLOP/s",
                          "frac": flops / 1e12 / (gemm_avg / 1e3) / peak,
                          "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained", "flops_per_launch": flops,
                          "avg_launch_ms": gemm_avg, "launches_timed": int(gemm_n), "traffic": None}
                         if B > 128 else
                         {"bound": "hbm", "kernel": "batch_gemm_pair_kernel<FILTER, %d>" % (64 if B <= 64 else 128),
                          "achieved": rows_per_gpu * DIM * 2 / 1e9 / (gemm_avg / 1e3),
                          "peak": float(peaks.get("hbm_gbs", 6650.0)), "unit": "GB/s",
                          "frac": rows_per_gpu * DIM * 2 / 1e9 / (gemm_avg / 1e3) / float(peaks.get("hbm_gbs", 6650.0)),
                          "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)",
                          "algorithmic_bytes_per_launch": rows_per_gpu * DIM * 2, "avg_launch_ms": gemm_avg,
                          "launches_timed": int(gemm_n), "traffic": None}),
            "clocks": sampler.summary() if sampler else None,
        })
    idx.close()
    dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def claim_stdout():
    """Keep a private handle on the real stdout for the ONE JSON line and point fd 1 at
    stderr, so banners printed by libraries (e.g. NCCL's version line) cannot pollute it."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from clip_database_b200 import GpuIndex
    from clip_database_b200.sharded import CudaShardBackend, ShardedIndex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (python -m torch.distributed.run --nproc-per-node %d ...)"
                             % (args.gpus, args.gpus))
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this benchmark has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    k = args.k
    if args.workload == "binary":
        return run_binary_workload(args, torch, device, local_rank)
    rows_per_gpu = args.rows or (10_000_000 if world == 1 else 12_500_000)
    rows = generate_rows(torch, device, rows_per_gpu, 1234 + rank, clustered=args.data == "clustered")
    idx = GpuIndex(local_rank)
    idx.attach(rows, rowid_base=1 + rank * rows_per_gpu)
    idx.set_option("scan_variant", args.variant)
    backend = CudaShardBackend(idx)          # puts the context on torch's current stream
    sharded = ShardedIndex(backend, fused={"auto": "auto", "fused": True, "nccl": False}[args.exchange]) if world > 1 else ShardedIndex(backend)
    args.exchange_used = ("fused peer-memory exchange in the scan kernel" if sharded.fused else
                          "NCCL all-gather + merge kernel") if world > 1 else "none (single GPU)"

    if args.workload == "batch" and world > 1:
        return run_sharded_batch_workload(args, torch, dist, idx, sharded, rows_per_gpu, device, rank, local_rank, world)
    if args.workload == "batch":
        return run_batch_workload(args, torch, idx, rows, rows_per_gpu, device, rank, local_rank)

    n_q = 64
    host_q = make_queries(rows, n_q, 99, args.data == "clustered")
    if args.workload == "blend":
        host_q2 = np.random.default_rng(100).standard_normal((n_q, DIM), dtype=np.float32)
        host_q2 /= np.linalg.norm(host_q2, axis=1, keepdims=True)
        host_neg = np.random.default_rng(101).standard_normal((n_q, DIM), dtype=np.float32)
        host_neg /= np.linalg.norm(host_neg, axis=1, keepdims=True)
        d_q2 = torch.from_numpy(host_q2).to(device)
        d_neg = torch.from_numpy(host_neg).to(device).view(n_q, 1, DIM)
        d_w = torch.tensor([[0.7, 0.3]] * n_q, dtype=torch.float32, device=device)
        d_nw = torch.full((n_q, 1), 0.5, dtype=torch.float32, device=device)
        d_blend = torch.empty((1, DIM), dtype=torch.float32, device=device)
    d_q = torch.from_numpy(host_q).to(device)

    o_ids = torch.empty((1, k), dtype=torch.int64, device=device)
    o_dist = torch.empty((1, k), dtype=torch.float32, device=device)
    o_n = torch.zeros(1, dtype=torch.int32, device=device)
    o_nan = torch.zeros(1, dtype=torch.int64, device=device)

    def search_dev(q_vec):
        if world == 1:
            idx.search_device(q_vec.view(1, -1), k, o_ids, o_dist, o_n, o_nan)
        else:
            sharded.search_device(q_vec, k)     # local scan + all-gather + merge

    def step_device(i):
        j = i % n_q
        if args.workload == "blend":
            idx.blend_device(d_q[j:j + 1], d_q2[j:j + 1], d_w[j:j + 1], d_neg[j:j + 1], d_nw[j:j + 1], d_blend)
            search_dev(d_blend[0])
        else:
            search_dev(d_q[j])

    def step_e2e(i):
        j = i % n_q
        if world == 1 and args.workload == "blend":
            return idx.blend_search(host_q[j], k, e2=host_q2[j], weights=(0.7, 0.3), negatives=[host_neg[j]],
                                    negative_weights=[0.5])
        if world == 1:
            return idx.search(host_q[j], k)
        return sharded.search(host_q[j], k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- value: device-resident query, K steps, CUDA events, max over ranks
    for i in range(args.warmup):
        step_device(i)
    barrier()
    if sampler:
        sampler.start()
    launches0 = idx.launch_count
    idx.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step_device(i)
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    scan_ms, scans = idx.profile_read()
    idx.profile(False)
    launches = idx.launch_count - launches0

    # ---- e2e: host query in, host results out, every step
    for i in range(min(args.warmup, 5)):
        step_e2e(i)
    barrier()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    ee0.record()
    last = None
    for i in range(args.steps):
        last = step_e2e(i)
    ee1.record()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t_wall) * 1e3
    e2e_ms = max_over_ranks(max(ee0.elapsed_time(ee1), 0.0))
    e2e_wall_ms = max_over_ranks(e2e_wall_ms)
    if sampler:
        sampler.stop()

    # sanity: the timed path produced a full, sorted answer
    if world == 1:
        assert last.counts[0] == k and np.all(np.diff(last.distances[0]) >= 0)
    else:
        assert len(last[0]) == k and np.all(np.diff(last[1]) >= 0)

    scan_ms_avg = scan_ms / max(scans, 1)
    scan_ms_avg = max_over_ranks(scan_ms_avg)
    total_rows = rows_per_gpu * world
    gb_per_query = total_rows * ROW_BYTES / 1e9
    ms_per_step = ms_total / args.steps
    value = gb_per_query / (ms_per_step / 1e3)
    e2e_ms_per_step = max(e2e_ms, e2e_wall_ms) / args.steps       # host-visible time bounds it
    e2e_value = gb_per_query / (e2e_ms_per_step / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback"
        achieved = rows_per_gpu * ROW_BYTES / 1e9 / (scan_ms_avg / 1e3)
        traffic = None       # dram read+write bytes per launch from the committed ncu --set full capture
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
            if int(cap.get("rows", 0)) == rows_per_gpu:      # only for the workload that was captured
                traffic = cap.get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic" if args.data == "uniform" else "synthetic, clustered (1024 clusters, queries near stored rows)",
            "config": workload_config(args, rows_per_gpu, world),
            "queries_per_s": 1e3 / ms_per_step,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ROW_BYTES * (3 if args.workload == "blend" else 1),
                    "d2h_bytes_per_step": k * 12 + 12, "ms_per_step": e2e_ms_per_step,
                    "queries_per_s": 1e3 / e2e_ms_per_step,
                    "api": "GpuIndex.search (clipdb_search)" if world == 1 else "ShardedIndex.search"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "scan_tma_kernel" if args.variant in (0, 1) else "scan_ldg_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8000": achieved / 8000.0, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": rows_per_gpu * ROW_BYTES,
                         "avg_launch_ms": scan_ms_avg, "launches_timed": int(scans), "traffic": traffic},
            "clocks": sampler.summary() if sampler else None,
        }
        if world == 1 and not args.no_cpu_baseline:
            sec, provider, extra = cpu_reference_path(args.cpu_sample_rows, args.cpu_sample_queries)
            cpu_gbps = args.cpu_sample_rows * ROW_BYTES / 1e9 / sec
            line["cpu_baseline"] = {
                "value": cpu_gbps, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": ("reference SQL statement (image_database.py:1564-1574) on SQLite via %s; %d-row sample, "
                           "%d queries, k=20, 1 thread; vec0 is a plain stand-in table"
                           % (provider, args.cpu_sample_rows, args.cpu_sample_queries)),
                "queries_per_s_extrapolated_to_config": 1.0 / (sec * total_rows / args.cpu_sample_rows),
                "cpu": cpu_model(), **extra}
        emit(line)

    idx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
