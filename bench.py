#!/usr/bin/env python3
"""bench.py — KNN scan throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

A "step" is one pass of the hot path over one query: scan every resident 1152-d row with
cosine distance and select the top k (k = 20).

  N = 1   BASELINE configs[1]: 10M float32 rows (46.08 GB) resident on one B200.  The same run
          also measures, on the same resident store and as sub-records of the ONE line,
          configs[3] (0.7/0.3 blend + negative), configs[2] (B = 256, k = 100 through the
          tcgen05 contraction + fp32 re-rank) and configs[0] (100k rows written to SQLite in the
          reference schema and loaded THROUGH THE LOADER, answered through ImageDatabase next to
          the reference's own SQL statement on that database), plus a `parity` block: the first
          100k rows against the CPU oracle and the full store against a float64 recomputation.
  N > 1   BASELINE configs[4]: 100M rows row-sharded over the N GPUs of the box (12.5M-row blocks,
          block b seeded 1234+b, so the data is the same for every N).  N = 8 and 4 hold their
          shard as float32 (57.6 / 115.2 GB per GPU); at N = 2 a shard is 230 GB of float32, so
          the store is bf16-primary: the bf16 copy is resident, every query is pre-selected by
          the tensor-core path and the exact float32 re-rank reads its candidates from a tiered
          float32 store (HBM + pinned host memory).  One launch per rank per query: the
          candidates cross NVLink inside the search kernel (peer memory, no collective call).
          The line carries a `parity` block (fused exchange == NCCL path == float64 recompute
          on every rank, with a planted cross-shard tie) and a `strong` sub-record (the 10M-row
          configs[1] store split over the N GPUs, with the in-kernel timeline).

`value` is whole-job scan GB/s with the query already in HBM (algorithmic bytes per query =
bytes of the store representation that is scanned); `e2e` is the same metric
through the public host API (host query in, host results out, copies inside the timed region).
Inputs (>= 46 GB per scan) are far larger than L2 (126 MB), so no flush is needed between steps.

The reference arm times the reference's statement (image_database.py:1564-1574) on the real
SQLite with the oracle's C restatement of sqlite-vec's vec_distance_cosine, single-threaded like
the reference, on a bounded row sample; the same thing is reported as `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
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
BF16_ROW_BYTES = DIM * 2
METRIC = "knn_scan_throughput"
UNIT = "GB/s"
RTOL = 1e-5                      # north_star: |delta| <= 1e-5 * max(|d|, 1)
BLOCK_ROWS = 12_500_000          # fixed generation blocks, block b <-> seed 1234 + b
SHARDED_TOTAL_ROWS = 100_000_000
CHUNK_ROWS = 500_000
CONFIG0_ROWS = 100_000
CONFIG0_QUERIES = 32


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--rows", type=int, default=0,
                    help="rows per GPU (default: 10M at N=1; 100M / N at N>1)")
    ap.add_argument("--workload", default="single", choices=["single", "blend", "batch", "binary"],
                    help="single: configs[1] (+ the sub-records, see --sub); blend: configs[3] alone; "
                         "batch: configs[2] alone (B queries per step through the tcgen05 contraction + fp32 re-rank); "
                         "binary: the sign-code fallback search (SURVEY §8 f-4) over bit-packed codes")
    ap.add_argument("--sub", default="all", choices=["all", "none"],
                    help="all: the default run also measures configs[0]/[2]/[3], parity and (N>1) strong scaling; "
                         "none: only the headline timing (profiling runs)")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--data", default="uniform", choices=["uniform", "clustered"],
                    help="uniform: seeded random unit rows and queries (BASELINE configs); clustered: 1024 clusters, "
                         "queries near stored rows (SURVEY §8d variant)")
    ap.add_argument("--sample-stride", type=int, default=0, help="batch: pass A samples 1/s of the rows (0 = auto)")
    ap.add_argument("--no-refine", action="store_true", help="batch: skip the second threshold")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fused", "nccl"],
                    help="N > 1: fused = the search kernel's last CTA exchanges candidates over NVLink peer memory "
                         "and merges (one launch per rank); nccl = all-gather + merge kernel; auto = fused when "
                         "every rank can map its peers' memory, else nccl")
    ap.add_argument("--store", default="auto", choices=["auto", "fp32", "bf16-primary"],
                    help="N > 1: how a shard is held (auto: float32 when it fits the GPU, else bf16-primary)")
    ap.add_argument("--variant", type=int, default=0, help="scan kernel: 0 auto, 1 TMA ring, 2 direct loads")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=CONFIG0_ROWS)
    ap.add_argument("--cpu-sample-queries", type=int, default=CONFIG0_QUERIES)
    return ap.parse_args()


# ---------------------------------------------------------------- host probes ---------------
def host_mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return None


def gpu_total_bytes(index=0):
    """Total memory of one GPU through NVML (no CUDA context: the reference arm calls this too)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        return int(pynvml.nvmlDeviceGetMemoryInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).total)
    except Exception:
        return None


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def committed_traffic(name, **match):
    """dram read+write bytes per launch of a kernel from a committed `ncu --set full` capture
    (profiles/<name>.json), or (None, None) when the capture was of another workload."""
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", name + ".json")))
    except Exception:
        return None, None
    for key, want in match.items():
        if int(cap.get(key, -1)) != int(want):
            return None, None
    return cap.get("dram_bytes_per_launch"), "replayed from %s (not measured in this run)" % cap.get("source", name)


# ---------------------------------------------------------------- the store plan ------------
def plan_store(args, n_gpus):
    """How many rows every GPU holds and in which representation.  A pure function of the
    arguments and of the box (GPU memory size, host memory), so both arms print the same config."""
    if n_gpus == 1:
        rows = args.rows or 10_000_000
        return {"rows_per_gpu": rows, "rows_total": rows, "store": "fp32", "hbm_fp32_rows": rows, "host_fp32_rows": 0,
                "blocks": "1 block, seed 1234"}
    rows = args.rows or SHARDED_TOTAL_ROWS // n_gpus
    hbm = gpu_total_bytes() or 191_000_000_000
    budget = hbm - 14_000_000_000          # context, workspaces, the generation chunk, fragmentation
    kind = args.store
    if kind == "auto":
        kind = "fp32" if rows * ROW_BYTES <= budget else "bf16-primary"
    plan = {"rows_per_gpu": rows, "rows_total": rows * n_gpus, "store": kind, "hbm_fp32_rows": rows, "host_fp32_rows": 0,
            "blocks": ("12.5M-row blocks, block b seeded 1234+b" if not args.rows else "1 block per rank, seed 1234+rank")}
    if kind == "bf16-primary":
        # bf16 copy resident (2304 B/row); the float32 rows (4608 B/row, read only by the re-rank) fill what
        # is left of HBM and continue in pinned host memory
        # MemAvailable moves by a few hundred MB between two looks; rounding it down to whole 8 GB keeps the plan
        # (and with it `config`) identical for the two arms of a run
        avail = (host_mem_available_bytes() or 200_000_000_000) // 8_000_000_000 * 8_000_000_000
        host_budget = int(avail * 0.80) // n_gpus

        def split(r):
            in_hbm = max(0, min(r, (budget - r * BF16_ROW_BYTES) // ROW_BYTES)) & ~127
            return in_hbm, r - in_hbm
        in_hbm, in_host = split(rows)
        if in_host * ROW_BYTES > host_budget or rows * BF16_ROW_BYTES > budget:
            # this box cannot hold the full shard: the largest one it can (whole 500k-row chunks)
            fit = min((host_budget + budget) // (ROW_BYTES + BF16_ROW_BYTES), budget // BF16_ROW_BYTES)
            rows = max(CHUNK_ROWS, fit // CHUNK_ROWS * CHUNK_ROWS)
            in_hbm, in_host = split(rows)
            plan["note"] = ("host memory of this box holds %d of the %d rows per GPU the 100M-row store needs"
                            % (rows, plan["rows_per_gpu"]))
        plan.update(rows_per_gpu=rows, rows_total=rows * n_gpus, hbm_fp32_rows=in_hbm, host_fp32_rows=in_host)
    return plan


def workload_config(args, plan, n_gpus):
    rows = plan["rows_per_gpu"]
    if n_gpus == 1:
        what = ("single-query cosine KNN, k=%d, %d x 1152 fp32 rows resident per GPU (BASELINE configs[1])"
                % (args.k, rows))
    else:
        what = ("row-sharded single-query cosine KNN, k=%d, %d x 1152 rows per GPU x %d GPUs = %d rows "
                "(BASELINE configs[4])" % (args.k, rows, n_gpus, plan["rows_total"]))
    cfg = {
        "workload": what,
        "rows_per_gpu": rows,
        "rows_total": plan["rows_total"],
        "dim": DIM,
        "k": args.k,
        "metric_kind": "cosine",
        "blend": args.workload == "blend",
        "store": ("float32 rows resident in HBM" if plan["store"] == "fp32" else
                  "bf16-primary: bf16 copy resident in HBM (tensor-core pre-selection), float32 rows for the exact "
                  "re-rank tiered: %d in HBM + %d in pinned host memory per GPU"
                  % (plan["hbm_fp32_rows"], plan["host_fp32_rows"])),
        "scanned_bytes_per_row": ROW_BYTES if plan["store"] == "fp32" else BF16_ROW_BYTES,
        "data_blocks": plan["blocks"],
        "l2": "inputs_larger_than_L2",
        "parallelism": "row-shard x%d" % n_gpus if n_gpus > 1 else "single GPU",
    }
    if "note" in plan:
        cfg["note"] = plan["note"]
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
def config0_database(tmp, n_rows, with_codes=True, cheap_rows=False):
    """BASELINE configs[0] / SURVEY §8d config 1: `default_rng(1234)` unit rows written to SQLite in the
    reference schema (vec0 stand-in table, image_embeddings, images, binary_embeddings).  ``cheap_rows``: the
    oracle's fast generator instead of numpy's (large timing-only samples)."""
    from clip_database_b200 import synth
    if cheap_rows:
        from oracle import ref
        rows = ref.fill_unit_rows(n_rows, DIM, 1234)
    else:
        rows = synth.unit_rows(n_rows, DIM, 1234)
    db_path = os.path.join(tmp, "config0_%d.db" % n_rows)
    synth.write_reference_db(db_path, rows, binary_codes=with_codes)
    return db_path, rows


def config0_queries(n):
    from clip_database_b200 import synth
    return synth.unit_rows(n, DIM, 99)


def time_reference_statement(db_path, queries, k, warmup):
    """The interval the reference itself calls `db_query` (image_database.py:1557-1631): its SQL statement
    executed by the real SQLite on one thread.  Returns (mean seconds, provider, per-query result rows)."""
    from oracle import sql_harness
    conn, provider = sql_harness.connect(db_path)
    for q in queries[:warmup]:
        sql_harness.run_statement(conn, q, k)
    times, results = [], []
    for q in queries[warmup:]:
        t0 = time.perf_counter()
        got = sql_harness.run_statement(conn, q, k, with_rowid=True)
        times.append(time.perf_counter() - t0)
        results.append(got)
    conn.close()
    return float(np.mean(times)), provider, results


def scalar_loop_rates(rows, queries):
    """The bare scalar loop + bounded top-k (no SQLite machinery): an upper bound on what the reference's
    single thread could reach, and the same loop on every host core."""
    from oracle import ref
    qs = queries[:8]
    t0 = time.perf_counter()
    for q in qs:
        ref.knn(rows, q, 20)
    loop_s = (time.perf_counter() - t0) / len(qs)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    for q in qs:
        ref.distances(rows, q, threads=cores)
    mt_s = (time.perf_counter() - t0) / len(qs)
    gb = rows.shape[0] * ROW_BYTES / 1e9
    return {"scalar_loop_1thread_GBps": gb / loop_s, "scalar_loop_all_cores_GBps": gb / mt_s, "host_cores": cores}


def cpu_baseline_record(sec, provider, sample_rows, n_queries, total_rows, extra):
    gbps = sample_rows * ROW_BYTES / 1e9 / sec
    return {"value": gbps, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": ("reference SQL statement (image_database.py:1564-1574) on SQLite via %s; %d-row sample of the "
                       "workload (the configs[0] database), %d timed queries, k=20, 1 thread (SQLite runs a statement on "
                       "one thread); vec0 is a plain stand-in table" % (provider, sample_rows, n_queries)),
            "sample_rows": sample_rows, "ms_per_query_at_sample": sec * 1e3,
            "queries_per_s_extrapolated_to_config": 1.0 / (sec * total_rows / sample_rows),
            "cpu": cpu_model(), **extra}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_gpus = args.gpus
    plan = plan_store(args, n_gpus)
    # bound the sample so warmup + steps statements end within about a minute
    per_row = 2.0e-6
    n_stmt = max(args.steps + args.warmup, 1)
    sample = int(min(args.cpu_sample_rows, max(10_000, 60.0 / n_stmt / per_row)))
    tmp = tempfile.mkdtemp(prefix="clipdb_cpu_")
    try:
        db_path, rows = config0_database(tmp, sample)
        warm = max(args.warmup, 1)
        queries = config0_queries(args.steps + warm)
        sec, provider, _ = time_reference_statement(db_path, queries, 20, warm)
        extra = scalar_loop_rates(rows, queries)
        os.remove(db_path)
        # a second, 10x larger sample when the run is short enough: the rate does not depend on the sample size
        larger = None
        big = sample * 10
        if n_stmt * big * per_row <= 100.0 and shutil.disk_usage(tmp).free > big * ROW_BYTES * 3:
            big_db, _big_rows = config0_database(tmp, big, with_codes=False, cheap_rows=True)
            del _big_rows
            nq = max(3, min(args.steps, 8))
            big_sec, _, _ = time_reference_statement(big_db, queries[:nq + 1], 20, 1)
            larger = {"sample_rows": big, "timed_queries": nq, "ms_per_query": big_sec * 1e3,
                      "value": big * ROW_BYTES / 1e9 / big_sec, "unit": UNIT}
            os.remove(big_db)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    cb = cpu_baseline_record(sec, provider, sample, args.steps, plan["rows_total"], extra)
    gbps = cb["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": gbps, "unit": UNIT, "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, plan, n_gpus),
        "sample_rows": sample, "larger_sample": larger,
        "queries_per_s_at_sample": 1.0 / sec,
        "queries_per_s_extrapolated_to_config": cb["queries_per_s_extrapolated_to_config"],
        "cpu_baseline": cb,
        "e2e": {"value": gbps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------- synthetic data ------------
N_CLUSTERS = 1024


def rank_blocks(args, plan, rank, n_gpus):
    """[(seed, rows)] generated by this rank, in scan order."""
    rows = plan["rows_per_gpu"]
    if n_gpus == 1 or args.rows:
        return [(1234 + rank, rows)]
    per = SHARDED_TOTAL_ROWS // n_gpus // BLOCK_ROWS          # blocks a full shard is made of
    first = rank * per
    out, left = [], rows
    for b in range(first, first + per):
        if left <= 0:
            break
        out.append((1234 + b, min(BLOCK_ROWS, left)))
        left -= BLOCK_ROWS
    return out


class Float64Truth:
    """Ground truth for a handful of queries while the rows stream by: every distance recomputed in
    float64 by torch (the task's "plain reference of the same op", in double), the best k + slack per
    query kept under the (distance, rowid) order."""

    def __init__(self, torch, device, queries, k, slack=16):
        self.t = torch
        self.q = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)).to(device).double()
        self.qn = self.q.norm(dim=1)
        self.keep = k + slack
        self.d = torch.empty((self.q.shape[0], 0), dtype=torch.float64, device=device)
        self.ids = torch.empty((self.q.shape[0], 0), dtype=torch.int64, device=device)

    def update(self, view, first_rowid):
        t = self.t
        for lo in range(0, view.shape[0], 250_000):
            r = view[lo:lo + 250_000].double()
            d = 1.0 - (self.q @ r.T) / (self.qn[:, None] * r.norm(dim=1)[None, :])
            m = min(self.keep, d.shape[1])
            vals, pos = t.topk(d, m, dim=1, largest=False, sorted=True)
            # topk may pick any of several equal distances at its cut: widen the cut to every row that ties it
            cut = vals[:, -1:]
            extra = (d == cut).sum(dim=1).max().item() if m < d.shape[1] else 0
            if extra > 1:
                vals, pos = t.topk(d, min(m + int(extra), d.shape[1]), dim=1, largest=False, sorted=True)
            self.d = t.cat([self.d, vals], dim=1)
            self.ids = t.cat([self.ids, pos + (first_rowid + lo)], dim=1)
            self._trim()

    def _trim(self):
        t = self.t
        order = t.argsort(self.ids, dim=1, stable=True)
        d, ids = t.gather(self.d, 1, order), t.gather(self.ids, 1, order)
        order = t.argsort(d, dim=1, stable=True)                  # stable: equal distances stay in rowid order
        self.d = t.gather(d, 1, order)[:, :self.keep].contiguous()
        self.ids = t.gather(ids, 1, order)[:, :self.keep].contiguous()

    def merge_from(self, others):
        """others: [(d numpy, ids numpy)] of the other shards."""
        t = self.t
        for d, ids in others:
            self.d = t.cat([self.d, t.from_numpy(d).to(self.d.device)], dim=1)
            self.ids = t.cat([self.ids, t.from_numpy(ids).to(self.ids.device)], dim=1)
        self._trim()

    def host(self):
        return self.d.cpu().numpy(), self.ids.cpu().numpy()


def compare_with_truth(got_ids, got_d, truth_d, truth_ids, k):
    """One query: (ids identical, id mismatches beyond a tolerance tie, max |d - d64|, distances within tolerance)."""
    k = min(k, len(got_ids))
    lookup = {int(i): float(d) for i, d in zip(truth_ids, truth_d)}
    identical = bool(np.array_equal(got_ids[:k], truth_ids[:k]))
    beyond, max_abs, dist_ok = 0, 0.0, True
    for j in range(k):
        own = lookup.get(int(got_ids[j]))
        exp = float(truth_d[j])
        tol = RTOL * max(abs(exp), 1.0)
        if own is None:                       # not even among the best k + slack
            beyond += 1
            dist_ok = False
            continue
        max_abs = max(max_abs, abs(float(got_d[j]) - own))
        if abs(float(got_d[j]) - own) > tol or abs(float(got_d[j]) - exp) > tol:
            dist_ok = False
        if int(got_ids[j]) != int(truth_ids[j]) and abs(own - exp) > tol:
            beyond += 1
    return identical, beyond, max_abs, dist_ok


def plant_vector():
    v = np.random.default_rng(777).standard_normal(DIM, dtype=np.float32)
    return v / np.linalg.norm(v)


def parity_queries(n_q):
    """Query 0 sits next to the planted rows (its best hits are exact cross-shard ties); the others are the
    bench's own seeded unit queries."""
    rng = np.random.default_rng(99)
    q = rng.standard_normal((n_q, DIM), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    near = plant_vector() + 0.05 * q[0]
    q[0] = near / np.linalg.norm(near)
    return np.ascontiguousarray(q, dtype=np.float32)


def generate_store(torch, device, blocks, first_rowid, clustered=False, into=None, commit=None, truth=None, plants=()):
    """randn float32 from a seeded CUDA generator, in 500k-row chunks, rows
    L2-normalised.  Written straight into ``into`` (a [n, 1152] CUDA tensor) or handed chunk by chunk to
    ``commit(chunk, position)`` (a store filled by appends).  ``plants``: positions that receive the plant
    vector (exact duplicates).  ``truth``: Float64Truth fed with every chunk.  ``clustered``: the §8d variant —
    row r = normalise(centre[r % 1024] + noise) with |noise| ~ 0.75 |centre|."""
    centres = None
    if clustered:
        cgen = torch.Generator(device=device)
        cgen.manual_seed(4242)                      # the same centres on every rank
        centres = torch.randn((N_CLUSTERS, DIM), generator=cgen, device=device)
        centres /= centres.norm(dim=1, keepdim=True)
    plant = torch.from_numpy(plant_vector()).to(device) if plants else None
    scratch = None if into is not None else torch.empty((CHUNK_ROWS, DIM), dtype=torch.float32, device=device)
    pos = 0
    for seed, n_rows in blocks:
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        for lo in range(0, n_rows, CHUNK_ROWS):
            m = min(CHUNK_ROWS, n_rows - lo)
            view = into[pos:pos + m] if into is not None else scratch[:m]
            view.normal_(generator=gen)
            if clustered:
                view.mul_(0.75 / DIM ** 0.5)
                view.add_(centres[torch.arange(pos, pos + m, device=device) % N_CLUSTERS])
            view.div_(view.norm(dim=1, keepdim=True))
            for p in plants:
                if pos <= p < pos + m:
                    view[p - pos] = plant
            if truth is not None:
                truth.update(view, first_rowid + pos)
            if commit is not None:
                commit(view, pos)
            pos += m
    return pos


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


def live_cublas_tflops(torch, device):
    """cuBLAS bf16 GEMM back to back for ~1 s right after a timed region: the same box in the same power
    state (the board sits at its power cap whenever the tensor pipe is busy, and the clock it settles at varies
    from box to box and minute to minute)."""
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
        return 2.0 * 8192 ** 3 * reps / 1e12 / (g0.elapsed_time(g1) / 1e3)
    except Exception:
        return None


# ---------------------------------------------------------------- configs[2]: batched -------
def measure_batch(args, torch, idx, rows_per_gpu, device, local_rank, host_q, steps, warmup, k, with_live=True,
                  check_exact=3):
    """B queries per step through the tcgen05 contraction + fp32 re-rank on an index whose bf16 store is enabled.
    Returns the record (metric queries/s; roofline of the filter-pass contraction kernel)."""
    n_sets, B = host_q.shape[0], host_q.shape[1]
    d_q = torch.from_numpy(host_q).to(device)
    o_ids = torch.empty((B, k), dtype=torch.int64, device=device)
    o_dist = torch.empty((B, k), dtype=torch.float32, device=device)
    o_n = torch.zeros(B, dtype=torch.int32, device=device)
    o_nan = torch.zeros(B, dtype=torch.int64, device=device)
    flags = torch.zeros(B, dtype=torch.int32, device=device)
    sampler = ClockSampler(local_rank)

    for i in range(warmup):
        idx.search_batch_device(d_q[i % n_sets], k, o_ids, o_dist, o_n, o_nan, flags)
    torch.cuda.synchronize()
    sampler.start()
    launches0 = idx.launch_count
    idx.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        idx.search_batch_device(d_q[i % n_sets], k, o_ids, o_dist, o_n, o_nan, flags)
    ev1.record()
    torch.cuda.synchronize()
    ms_step = ev0.elapsed_time(ev1) / steps
    gemm_ms, gemm_n = idx.profile_read()
    kernel_mhz, _ = idx.profile_clock()
    idx.profile(False)
    launches = idx.launch_count - launches0
    flagged = int((flags != 0).sum())
    cand, surv = idx.batch_stats()

    for i in range(min(warmup, 3)):
        idx.search(host_q[i % n_sets], k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        last = idx.search(host_q[i % n_sets], k)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    sampler.stop()
    assert np.all(last.counts == k)

    # parity: the batched answers against the exact single-query path (same context, batch path bypassed)
    exact_checked = exact_same = 0
    if check_exact:
        old_min = idx.get_option("batch_min_nq")
        idx.set_option("batch_min_nq", 1 << 20)
        picks = sorted(set(np.linspace(0, B - 1, check_exact).astype(int).tolist()))
        for j in picks:
            one = idx.search(host_q[(steps - 1) % n_sets][j], k)
            exact_checked += 1
            exact_same += int(np.array_equal(one.rowids[0], last.rowids[j]) and
                              np.array_equal(one.distances[0].view(np.uint32), last.distances[j].view(np.uint32)))
        idx.set_option("batch_min_nq", old_min)
        if exact_same != exact_checked:
            raise SystemExit("PARITY FAILURE: the batched path and the exact scan disagree")

    peaks = load_peaks()
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flops = 2.0 * B * rows_per_gpu * DIM
    live = live_cublas_tflops(torch, device) if with_live else None
    gemm_avg = gemm_ms / max(gemm_n, 1)
    tflops = flops / 1e12 / (gemm_avg / 1e3)
    store_gbps = rows_per_gpu * BF16_ROW_BYTES / 1e9 / (gemm_avg / 1e3)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    if B > 128:     # 256 queries per pass: AI = 256 flop/B, above the ridge -> tensor-bound
        traffic, traffic_src = committed_traffic("batch_traffic", rows=rows_per_gpu, batch=B)
        roof = {"bound": "tensor", "kernel": "batch_gemm_pair_kernel<FILTER, 256>", "achieved": tflops, "peak": peak,
                "unit": "TFLOP/s", "frac": tflops / peak, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained",
                "flops_per_launch": flops, "avg_launch_ms": gemm_avg, "launches_timed": int(gemm_n),
                "hbm_GBps_bf16_store": store_gbps, "hbm_frac_bf16_store": store_gbps / hbm_peak,
                "live_cublas_tflops": live, "frac_of_live_cublas": (tflops / live) if live else None,
                # the clock the launches really ran at (%clock64 / %globaltimer inside the kernel): the board holds
                # its power cap by lowering it, and the tensor pipe's ceiling moves with it (nominal dense bf16 =
                # 2250 TFLOP/s at 1965 MHz)
                "sm_mhz_in_kernel": kernel_mhz,
                "nominal_peak_at_that_clock": (2250.0 * kernel_mhz / 1965.0) if kernel_mhz else None,
                "frac_of_peak_at_that_clock": (tflops / (2250.0 * kernel_mhz / 1965.0)) if kernel_mhz else None,
                "traffic": traffic, "traffic_source": traffic_src}
    else:           # 64 / 128 queries per pass: the contraction streams the bf16 store -> HBM-bound
        roof = {"bound": "hbm", "kernel": "batch_gemm_pair_kernel<FILTER, %d>" % (64 if B <= 64 else 128),
                "achieved": store_gbps, "peak": hbm_peak, "unit": "GB/s", "frac": store_gbps / hbm_peak,
                "frac_of_nominal_8000": store_gbps / 8000.0, "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)",
                "algorithmic_bytes_per_launch": rows_per_gpu * BF16_ROW_BYTES, "avg_launch_ms": gemm_avg,
                "launches_timed": int(gemm_n), "tensor_TFLOPs": tflops, "sm_mhz_in_kernel": kernel_mhz,
                "traffic": None, "traffic_source": None}
    return {
        "metric": "knn_batched_queries_per_s", "value": B * 1e3 / ms_step, "unit": "queries/s", "n_gpus": 1,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16 pre-select + f32 re-rank",
        "data": "synthetic" if args.data == "uniform" else "synthetic, clustered (1024 clusters, queries near stored rows)",
        "config": {"workload": "batched cosine KNN, B=%d queries per step, k=%d, %d x 1152 rows "
                               "(BASELINE configs[2]): tcgen05 contraction + fp32 re-rank" % (B, k, rows_per_gpu),
                   "rows_per_gpu": rows_per_gpu, "batch": B, "k": k, "dim": DIM,
                   "store_bytes": {"fp32": rows_per_gpu * ROW_BYTES, "bf16": rows_per_gpu * BF16_ROW_BYTES},
                   "l2": "inputs_larger_than_L2", "sample_stride": args.sample_stride,
                   "refine": not args.no_refine},
        "equivalent_scan_GBps_fp32": B * rows_per_gpu * ROW_BYTES / 1e9 / (ms_step / 1e3),
        "e2e": {"value": B * 1e3 / e2e_ms, "unit": "queries/s", "h2d_bytes_per_step": B * ROW_BYTES,
                "d2h_bytes_per_step": B * (k * 12 + 12) + B * 4, "ms_per_step": e2e_ms,
                "api": "GpuIndex.search (clipdb_search, nq=%d)" % B},
        "gpu_launches": int(launches), "flagged_queries": flagged,
        "exact_path_spot_check": {"queries": exact_checked, "bit_identical": exact_same},
        "candidates_per_query": {"filter_mean": float(cand[:B].mean()), "filter_max": int(cand[:B].max()),
                                 "reranked_mean": float(surv[:B].mean()), "reranked_max": int(surv[:B].max())},
        "roofline": roof,
        "clocks": sampler.summary(),
    }


def run_batch_workload(args, torch, idx, rows, rows_per_gpu, device, local_rank):
    """`--workload batch`: BASELINE configs[2] on its own (pass --k 100)."""
    B, k = args.batch, args.k
    idx.enable_batch()
    idx.set_option("batch_sample_stride", args.sample_stride)
    idx.set_option("batch_refine", 0 if args.no_refine else 1)
    idx.set_option("batch_min_nq", 1)          # --batch 1: one query through the bf16 pre-selection
    host_q = make_queries(rows, 4 * B, 99, args.data == "clustered").reshape(4, B, DIM)
    emit(measure_batch(args, torch, idx, rows_per_gpu, device, local_rank, host_q, args.steps, args.warmup, k))
    idx.close()
    return 0


# ---------------------------------------------------------------- sign codes ----------------
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
    peak = float(load_peaks().get("hbm_gbs", 6650.0))
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
                     "launches_timed": int(scans), "traffic": None, "traffic_source": None},
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
        shutil.rmtree(tmp, ignore_errors=True)
        line["cpu_baseline"] = {"value": m * (DIM // 8) / 1e9 / sec, "unit": "GB/s", "cores": 1, "kind": "port",
                                "rows_per_s": m / sec,
                                "sample": "literal restatement of image_database.py:1591-1629 (real SQLite fetchall + "
                                          "np.frombuffer + np.dot per row, 1 thread), %d-row sample, 3 queries; GB/s in "
                                          "the same packed-store bytes (the reference itself reads 1152 B per row)" % m,
                                "cpu": cpu_model()}
    emit(line)
    idx.close()
    return 0


# ---------------------------------------------------------------- configs[0] ----------------
def measure_config0(args, local_rank):
    """BASELINE configs[0]: 100k rows in the reference's SQLite schema, loaded into HBM THROUGH THE LOADER and
    answered through ``ImageDatabase.search_embedding`` (the repo's search path), one query at a time, k = 20 —
    next to the reference's own SQL statement executed by the real SQLite on the same database.  Returns
    (record, cpu_baseline record)."""
    from clip_database_b200 import ImageDatabase
    n, k = args.cpu_sample_rows, 20
    nq = max(args.cpu_sample_queries, 4)
    tmp = tempfile.mkdtemp(prefix="clipdb_cfg0_")
    try:
        db_path, rows = config0_database(tmp, n)
        queries = config0_queries(nq + 2)
        t0 = time.perf_counter()
        db = ImageDatabase(db_path, device=local_rank)
        open_s = time.perf_counter() - t0          # context creation + load
        load_s = db.load_seconds                    # the load itself (ImageDatabase.reload)
        load_source = db.load_source
        for q in queries[:2]:
            db.search_embedding(q, k=k, show_duplicates=True)
        t0 = time.perf_counter()
        got = [db.search_embedding(q, k=k, show_duplicates=True) for q in queries[2:]]
        gpu_s = (time.perf_counter() - t0) / nq
        launches = db.index.launch_count
        db.close()
        sec, provider, ref_rows = time_reference_statement(db_path, queries, k, 2)
        # parity: (file_path, similarity) lists of the GPU path against the reference statement's rows
        identical = beyond = 0
        max_abs = 0.0
        for mine, theirs in zip(got, ref_rows):
            paths_ref = [r[0] for r in theirs]
            sims_ref = [1.0 - r[2] for r in theirs]
            paths_mine = [p for p, _ in mine]
            identical += int(paths_mine == paths_ref)
            if len(mine) != len(theirs):
                beyond += abs(len(mine) - len(theirs))
            by_path = dict(zip(paths_ref, sims_ref))
            for j, (p, s) in enumerate(mine[:len(theirs)]):
                tol = RTOL * max(abs(1.0 - sims_ref[j]), 1.0)
                max_abs = max(max_abs, abs(s - sims_ref[j]))
                if p != paths_ref[j] and (p not in by_path or abs(by_path[p] - sims_ref[j]) > tol):
                    beyond += 1
                if abs(s - sims_ref[j]) > tol:
                    beyond += 1
        extra = scalar_loop_rates(rows, queries)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    ok = beyond == 0
    rec = {
        "workload": "BASELINE configs[0]: single-query cosine KNN, k=20 over %d x 1152 fp32 rows stored in SQLite "
                    "(reference schema), loaded through the streaming loader (ImageDatabase.reload: native SQLite "
                    "reader -> pinned buffer -> append), searched through ImageDatabase.search_embedding" % n,
        "rows": n, "queries": nq, "k": k,
        "loader_rows_per_s": n / load_s, "loader_MBps": n * ROW_BYTES / 1e6 / load_s, "load_s": load_s,
        "loader": load_source, "open_s": open_s,
        "ms_per_query_e2e": gpu_s * 1e3, "queries_per_s_e2e": 1.0 / gpu_s,
        "scan_GBps_e2e": n * ROW_BYTES / 1e9 / gpu_s,
        "reference_ms_per_query": sec * 1e3, "reference_provider": provider,
        "speedup_vs_reference_statement": sec / gpu_s,
        "gpu_launches": int(launches),
        "parity": {"against": "the reference's SQL statement on the same database (real SQLite + vec_shim)",
                   "queries": nq, "lists_identical": identical, "mismatches_beyond_tie": beyond,
                   "max_abs_similarity_diff": max_abs, "ok": ok},
    }
    return rec, cpu_baseline_record(sec, provider, n, nq, 10_000_000, extra)


# ---------------------------------------------------------------- N = 1 ---------------------
def timed_loop(torch, steps, fn):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        fn(i)
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / steps


def run_single_gpu(args, torch, device, local_rank):
    from clip_database_b200 import GpuIndex
    from oracle import blend as oblend
    from oracle import ref

    k = args.k
    plan = plan_store(args, 1)
    n = plan["rows_per_gpu"]
    subs = args.sub == "all" and args.workload == "single" and args.data == "uniform"
    n_parity = 4
    pq = parity_queries(8)
    truth = Float64Truth(torch, device, pq[:n_parity], k) if subs else None
    plants = (7, n // 2, n - 1) if subs and n > 16 else ()
    rows = torch.empty((n, DIM), dtype=torch.float32, device=device)
    generate_store(torch, device, rank_blocks(args, plan, 0, 1), 1, clustered=args.data == "clustered", into=rows,
                   truth=truth, plants=plants)
    idx = GpuIndex(local_rank)
    idx.attach(rows, rowid_base=1)
    idx.set_option("scan_variant", args.variant)
    idx.use_torch_stream()

    if args.workload == "batch":
        return run_batch_workload(args, torch, idx, rows, n, device, local_rank)

    # ---- parity, before anything is timed ------------------------------------------------------
    parity = None
    if subs:
        parity = {"tolerance": "|d - d_ref| <= 1e-5 * max(|d_ref|, 1); ids identical except where the reference "
                               "distances tie within that tolerance"}
        # (a) first 100k rows against the CPU oracle (C restatement of the scalar float32 loop + SQLite's top-k)
        m = min(CONFIG0_ROWS, n)
        head = GpuIndex(local_rank)
        head.attach(rows[:m], rowid_base=1)
        host_head = rows[:m].cpu().numpy()
        ident = beyond = 0
        max_abs = 0.0
        ok = True
        for q in pq:
            res = head.search(q, k)
            o_ids, o_d, _, o_nan = ref.knn(host_head, q, k, rowids=np.arange(1, m + 1))
            all_d = ref.distances(host_head, q)
            got_ids, got_d = res.row(0)
            same = got_ids == o_ids
            tol = RTOL * np.maximum(np.abs(o_d.astype(np.float64)), 1.0)
            max_abs = max(max_abs, float(np.abs(got_d.astype(np.float64) - o_d).max()))
            ok &= bool(np.all(np.abs(got_d.astype(np.float64) - o_d) <= tol)) and int(res.nan_rows[0]) == o_nan
            ident += int(same.all())
            beyond += int((np.abs(all_d[got_ids[~same] - 1].astype(np.float64) - o_d[~same]) > tol[~same]).sum())
        head.close()
        parity["first_100k_vs_oracle"] = {"rows": m, "queries": len(pq), "k": k, "lists_identical": ident,
                                          "id_mismatches_beyond_tie": beyond, "max_abs_d": max_abs,
                                          "ok": bool(ok and beyond == 0)}
        # (b) the whole store against the float64 recomputation made while the rows were generated
        t_d, t_ids = truth.host()
        res = idx.search(pq[:n_parity], k)
        ident = beyond = 0
        max_abs = 0.0
        ok = True
        for j in range(n_parity):
            got_ids, got_d = res.row(j)
            same, b, ma, dok = compare_with_truth(got_ids, got_d, t_d[j], t_ids[j], k)
            ident += int(same)
            beyond += b
            max_abs = max(max_abs, ma)
            ok &= dok
        tie_ids = res.rowids[0, :len(plants)].tolist()
        tie_ok = tie_ids == sorted(p + 1 for p in plants) and len(set(res.distances[0, :len(plants)].tolist())) == 1
        parity["full_store_vs_float64"] = {"rows": n, "queries": n_parity, "k": k, "lists_identical": ident,
                                           "id_mismatches_beyond_tie": beyond, "max_abs_d": max_abs,
                                           "planted_exact_ties_in_rowid_order": bool(tie_ok),
                                           "ok": bool(ok and beyond == 0 and tie_ok)}
        del truth
        if not (parity["first_100k_vs_oracle"]["ok"] and parity["full_store_vs_float64"]["ok"]):
            emit({"metric": METRIC, "error": "PARITY FAILURE", "parity": parity})
            return 3

    n_q = 64
    host_q = make_queries(rows, n_q, 99, args.data == "clustered")
    host_q2 = np.random.default_rng(100).standard_normal((n_q, DIM), dtype=np.float32)
    host_q2 /= np.linalg.norm(host_q2, axis=1, keepdims=True)
    host_neg = np.random.default_rng(101).standard_normal((n_q, DIM), dtype=np.float32)
    host_neg /= np.linalg.norm(host_neg, axis=1, keepdims=True)
    d_q = torch.from_numpy(host_q).to(device)
    d_q2 = torch.from_numpy(host_q2).to(device)
    d_neg = torch.from_numpy(host_neg).to(device).view(n_q, 1, DIM)
    d_w = torch.tensor([[0.7, 0.3]] * n_q, dtype=torch.float32, device=device)
    d_nw = torch.full((n_q, 1), 0.5, dtype=torch.float32, device=device)
    d_blend = torch.empty((1, DIM), dtype=torch.float32, device=device)
    o_ids = torch.empty((1, k), dtype=torch.int64, device=device)
    o_dist = torch.empty((1, k), dtype=torch.float32, device=device)
    o_n = torch.zeros(1, dtype=torch.int32, device=device)
    o_nan = torch.zeros(1, dtype=torch.int64, device=device)

    def step_plain(i):
        idx.search_device(d_q[i % n_q].view(1, -1), k, o_ids, o_dist, o_n, o_nan)

    def step_blend(i):
        j = i % n_q
        idx.blend_device(d_q[j:j + 1], d_q2[j:j + 1], d_w[j:j + 1], d_neg[j:j + 1], d_nw[j:j + 1], d_blend)
        idx.search_device(d_blend, k, o_ids, o_dist, o_n, o_nan)

    def e2e_plain(i):
        return idx.search(host_q[i % n_q], k)

    def e2e_blend(i):
        j = i % n_q
        return idx.blend_search(host_q[j], k, e2=host_q2[j], weights=(0.7, 0.3), negatives=[host_neg[j]],
                                negative_weights=[0.5])

    def measure(step_dev, step_host, steps, warmup):
        """(device ms/step, scan-kernel ms/launch, launches timed, kernel launches, e2e ms/step, last e2e result)"""
        for i in range(warmup):
            step_dev(i)
        torch.cuda.synchronize()
        launches0 = idx.launch_count
        idx.profile(True)
        ms = timed_loop(torch, steps, step_dev)
        scan_ms, scans = idx.profile_read()
        idx.profile(False)
        launches = idx.launch_count - launches0
        for i in range(min(warmup, 5)):
            step_host(i)
        torch.cuda.synchronize()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        ee0.record()
        last = None
        for i in range(steps):
            last = step_host(i)
        ee1.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t_wall) * 1e3
        e2e = max(ee0.elapsed_time(ee1), wall_ms) / steps          # host-visible time bounds it
        assert last.counts[0] == k and np.all(np.diff(last.distances[0]) >= 0)
        return ms, scan_ms / max(scans, 1), int(scans), int(launches), e2e, last

    blend_main = args.workload == "blend"
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step, scan_avg, scans, launches, e2e_ms, _ = measure(step_blend if blend_main else step_plain,
                                                            e2e_blend if blend_main else e2e_plain,
                                                            args.steps, args.warmup)
    sampler.stop()

    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback"
    gb = n * ROW_BYTES / 1e9
    traffic, traffic_src = committed_traffic("scan_traffic", rows=n)

    def roofline(avg_ms, launches_timed):
        achieved = gb / (avg_ms / 1e3)
        return {"bound": "hbm", "kernel": "scan_tma_kernel" if args.variant in (0, 1) else "scan_ldg_kernel",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8000": achieved / 8000.0, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": n * ROW_BYTES, "avg_launch_ms": avg_ms,
                "launches_timed": launches_timed, "traffic": traffic, "traffic_source": traffic_src}

    line = {
        "metric": METRIC, "value": gb / (ms_step / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic" if args.data == "uniform" else "synthetic, clustered (1024 clusters, queries near stored rows)",
        "config": workload_config(args, plan, 1),
        "queries_per_s": 1e3 / ms_step,
        "e2e": {"value": gb / (e2e_ms / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": ROW_BYTES * (3 if blend_main else 1),
                "d2h_bytes_per_step": k * 12 + 12, "ms_per_step": e2e_ms, "queries_per_s": 1e3 / e2e_ms,
                "api": "GpuIndex.blend_search (clipdb_blend_search)" if blend_main else "GpuIndex.search (clipdb_search)"},
        "gpu_launches": launches,
        "exchange": "none (single GPU)",
        "roofline": roofline(scan_avg, scans),
        "clocks": sampler.summary(),
    }

    if subs:
        line["parity"] = parity
        sub_steps = max(5, min(args.steps, 20))
        sub_warm = max(3, min(args.warmup, 5))
        # ---- configs[3]: text+image blend (0.7/0.3) with a negative prompt, then the same scan
        b_ms, b_scan, b_scans, b_launch, b_e2e, b_last = measure(step_blend, e2e_blend, sub_steps, sub_warm)
        blend_only = timed_loop(torch, 200, lambda i: idx.blend_device(
            d_q[i % n_q:i % n_q + 1], d_q2[i % n_q:i % n_q + 1], d_w[:1], d_neg[i % n_q:i % n_q + 1], d_nw[:1], d_blend))
        # the blended query against the oracle's numpy restatement of image_database.py:1378-1398 / 545-571
        q_ok = 0
        for j in range(4):
            _, q_gpu, flags = idx.blend_search(host_q[j], k, e2=host_q2[j], weights=(0.7, 0.3), negatives=[host_neg[j]],
                                               negative_weights=[0.5], return_query=True)
            q_ref = oblend.compose_query(host_q[j], host_q2[j], (0.7, 0.3), [host_neg[j]], [0.5])
            q_ok += int(flags == 0 and np.all(np.abs(q_gpu - q_ref) <= 1e-6 * np.abs(q_ref) + 3e-7 * np.abs(q_ref).max()))
        line["configs3"] = {
            "workload": "BASELINE configs[3]: weighted blend (0.7/0.3) + one negative (0.5), normalise, then the "
                        "k=20 scan over the same %d rows" % n,
            "steps": sub_steps, "warmup": sub_warm, "ms_per_step": b_ms, "value": gb / (b_ms / 1e3), "unit": UNIT,
            "queries_per_s": 1e3 / b_ms, "blend_kernel_us": blend_only * 1e3, "gpu_launches": b_launch,
            "e2e": {"value": gb / (b_e2e / 1e3), "unit": UNIT, "ms_per_step": b_e2e, "h2d_bytes_per_step": 3 * ROW_BYTES,
                    "d2h_bytes_per_step": k * 12 + 12, "api": "GpuIndex.blend_search (clipdb_blend_search)"},
            "roofline": roofline(b_scan, b_scans),
            "blended_query_matches_oracle": "%d/4" % q_ok,
        }
        if q_ok != 4:
            emit({"metric": METRIC, "error": "PARITY FAILURE (blend)", "configs3": line["configs3"]})
            return 3
        # ---- configs[2]: B = 256, k = 100 through the tcgen05 contraction + fp32 re-rank (same store + bf16 copy)
        t0 = time.perf_counter()
        idx.enable_batch()
        build_s = time.perf_counter() - t0
        idx.set_option("batch_min_nq", 1)
        B = args.batch
        bq = make_queries(rows, 4 * B, 99, False).reshape(4, B, DIM)
        rec = measure_batch(args, torch, idx, n, device, local_rank, bq, sub_steps, sub_warm, 100, check_exact=8)
        rec["bf16_store_build_s"] = build_s
        # one query at a time through the same pre-selection (what ImageDatabase(batch_store=True) does)
        one = measure_batch(args, torch, idx, n, device, local_rank, bq[:, :1], sub_steps, sub_warm, k, with_live=False,
                            check_exact=1)
        rec["single_query_through_bf16_preselect"] = {
            "ms_per_step": one["ms_per_step"], "queries_per_s": one["value"], "k": k,
            "hbm_GBps_bf16_store": one["roofline"]["achieved"], "frac_of_hbm_peak": one["roofline"]["frac"],
            "sm_mhz_in_kernel": one["roofline"].get("sm_mhz_in_kernel"),
            "e2e_ms_per_step": one["e2e"]["ms_per_step"], "gpu_launches": one["gpu_launches"]}
        idx.set_option("batch_min_nq", 2)
        line["configs2"] = rec
    idx.close()
    del rows
    torch.cuda.empty_cache()

    if subs and not args.no_cpu_baseline:
        rec0, cb = measure_config0(args, local_rank)
        line["configs0"] = rec0
        line["cpu_baseline"] = cb
        if not rec0["parity"]["ok"]:
            emit({"metric": METRIC, "error": "PARITY FAILURE (configs0)", "configs0": rec0})
            return 3
    elif not args.no_cpu_baseline:
        tmp = tempfile.mkdtemp(prefix="clipdb_cpu_")
        try:
            db_path, host_rows = config0_database(tmp, args.cpu_sample_rows)
            queries = config0_queries(args.cpu_sample_queries + 2)
            sec, provider, _ = time_reference_statement(db_path, queries, 20, 2)
            line["cpu_baseline"] = cpu_baseline_record(sec, provider, args.cpu_sample_rows, args.cpu_sample_queries,
                                                       n, scalar_loop_rates(host_rows, queries))
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    emit(line)
    return 0


# ---------------------------------------------------------------- N > 1 ---------------------
def run_sharded(args, torch, dist, device, rank, local_rank, world):
    from clip_database_b200 import GpuIndex
    from clip_database_b200.sharded import CudaShardBackend, ShardedIndex

    k = args.k
    plan = plan_store(args, world)
    n = plan["rows_per_gpu"]
    subs = args.sub == "all" and args.workload == "single" and args.data == "uniform"
    fused_arg = {"auto": "auto", "fused": True, "nccl": False}[args.exchange]

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def gather_objects(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def fail(what, detail):
        if rank == 0:
            emit({"metric": METRIC, "n_gpus": world, "error": "PARITY FAILURE: " + what, "detail": detail})
        barrier()
        dist.destroy_process_group()
        return 3

    strong = None
    if subs:
        strong = measure_strong_scaling(args, torch, dist, device, rank, local_rank, world, barrier, max_over_ranks,
                                        gather_objects)
        if strong.get("error"):
            return fail("strong-scaling store", strong)

    # ---- the sharded store -----------------------------------------------------------------
    first_rowid = 1 + rank * n
    n_parity = 4
    pq = parity_queries(8)
    truth = Float64Truth(torch, device, pq[:n_parity], k) if subs else None
    plants = (7, n // 2, n - 1) if subs else ()          # the same vector in every shard: cross-shard exact ties
    idx = GpuIndex(local_rank)
    t_load = time.perf_counter()
    if plan["store"] == "fp32":
        rows = torch.empty((n, DIM), dtype=torch.float32, device=device)
        generate_store(torch, device, rank_blocks(args, plan, rank, world), first_rowid, args.data == "clustered",
                       into=rows, truth=truth, plants=plants)
        idx.attach(rows, rowid_base=first_rowid)
    else:
        rows = None
        # the tiers are sized from what the box reports; if it then refuses the allocation anyway (a cgroup limit,
        # locked-memory limits, fragmentation) every rank falls back to a smaller shard together, and says so
        while True:
            try:
                idx.reserve(n, DIM, explicit_rowids=True, placement="host", device_rows=plan["hbm_fp32_rows"])
                idx.enable_batch()              # every append converts its rows into the bf16 copy
                ok = True
            except Exception as e:              # noqa: BLE001
                ok = False
                why = str(e)
            if all(gather_objects(ok)):
                break
            idx.close()
            idx = GpuIndex(local_rank)
            torch.cuda.empty_cache()
            if n <= 4 * CHUNK_ROWS:
                return fail("store allocation", {"rank": rank, "error": why if not ok else "another rank failed"})
            n = max(4 * CHUNK_ROWS, int(n * 0.8) // CHUNK_ROWS * CHUNK_ROWS)
            shrink = plan["rows_per_gpu"] - n
            plan = dict(plan, rows_per_gpu=n, rows_total=n * world,
                        host_fp32_rows=max(0, plan["host_fp32_rows"] - shrink),
                        note="the box refused the planned allocation: shard reduced to %d rows per GPU" % n)
            plan["hbm_fp32_rows"] = (n - plan["host_fp32_rows"]) & ~127
            plan["host_fp32_rows"] = n - plan["hbm_fp32_rows"]
            first_rowid = 1 + rank * n
            plants = (7, n // 2, n - 1) if subs else ()

        def commit(chunk, pos):
            ids = torch.arange(first_rowid + pos, first_rowid + pos + chunk.shape[0], dtype=torch.int64, device=device)
            idx.append(chunk, ids)
        generate_store(torch, device, rank_blocks(args, plan, rank, world), first_rowid, args.data == "clustered",
                       commit=commit, truth=truth, plants=plants)
        idx.set_option("batch_min_nq", 1)
        torch.cuda.empty_cache()
    load_s = time.perf_counter() - t_load
    idx.set_option("scan_variant", args.variant)
    bf16_primary = plan["store"] != "fp32"

    backend = CudaShardBackend(idx)              # puts the context on torch's current stream
    sharded = ShardedIndex(backend, fused=fused_arg)
    exchange_used = ("fused peer-memory exchange inside the search kernel (one launch per rank, no collective)"
                     if sharded.fused else "NCCL all-gather of k candidates per rank + merge kernel")
    nccl_side = ShardedIndex(CudaShardBackend(idx), fused=False) if subs else None

    def search_via(sh, q_host):
        """One query through a ShardedIndex, by the path the store's representation dictates."""
        if bf16_primary:
            ids, d, cnt = sh.search_batch(q_host[None, :], k)
            return ids[0, :cnt[0]].copy(), d[0, :cnt[0]].copy()
        return sh.search(q_host, k)

    # ---- parity over the sharded store, on every rank -----------------------------------------
    parity = None
    if subs:
        mine = truth.host()
        truth.merge_from([o for r, o in enumerate(gather_objects(mine)) if r != rank])
        t_d, t_ids = truth.host()
        del truth
        ident = beyond = 0
        max_abs = 0.0
        ok = True
        answers = []
        paths_agree = True
        for j in range(n_parity):
            f_ids, f_d = search_via(sharded, pq[j])
            n_ids, n_d = search_via(nccl_side, pq[j])
            answers.append((f_ids.tolist(), f_d.view(np.uint32).tolist()))
            paths_agree &= bool(np.array_equal(f_ids, n_ids) and np.array_equal(f_d.view(np.uint32), n_d.view(np.uint32)))
            same, b, ma, dok = compare_with_truth(f_ids, f_d, t_d[j], t_ids[j], k)
            ident += int(same)
            beyond += b
            max_abs = max(max_abs, ma)
            ok &= dok
        # the planted rows: `world` x 3 exact duplicates, one set per shard, must lead query 0 in rowid order
        want_tie = sorted(r * n + p + 1 for r in range(world) for p in plants)[:k]
        tie_ok = answers[0][0][:len(want_tie)] == want_tie and len(set(answers[0][1][:len(want_tie)])) == 1
        everyone = gather_objects((answers, paths_agree, ok, beyond))
        all_same = all(a[0] == everyone[0][0] for a in everyone)
        paths_agree = all(a[1] for a in everyone)
        ok = all(a[2] for a in everyone)
        beyond = max(a[3] for a in everyone)
        exact_tiered = None
        if bf16_primary:
            # the exact scan over the tiered float32 store (its host tier streams over PCIe: once is enough)
            e_ids, e_d = sharded.search(pq[1], k)
            exact_tiered = bool(np.array_equal(e_ids, np.asarray(answers[1][0])) and
                                np.array_equal(e_d.view(np.uint32), np.asarray(answers[1][1], dtype=np.uint32)))
            exact_tiered = all(gather_objects(exact_tiered))
        parity = {"tolerance": "|d - d_ref| <= 1e-5 * max(|d_ref|, 1); ids identical except where float64 distances "
                               "tie within that tolerance",
                  "queries": n_parity, "k": k, "rows_total": plan["rows_total"],
                  "reference": "float64 torch recomputation of every distance on every shard, merged by (distance, rowid)",
                  "lists_identical": ident, "id_mismatches_beyond_tie": beyond, "max_abs_d": max_abs,
                  "fused_equals_nccl_bitwise": bool(paths_agree) if sharded.fused else None,
                  "all_ranks_hold_identical_answers": bool(all_same),
                  "planted_cross_shard_ties_in_rowid_order": bool(tie_ok),
                  "tie_rows": len(want_tie), "exact_scan_over_tiered_store_equals_preselect": exact_tiered,
                  "ok": bool(ok and beyond == 0 and all_same and paths_agree and tie_ok and exact_tiered is not False)}
        if not parity["ok"]:
            return fail("sharded answers", parity)

    # ---- the timed region ---------------------------------------------------------------------
    n_q = 64
    if bf16_primary and args.data == "clustered":
        raise SystemExit("--data clustered needs a float32-resident store")
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

    def search_dev(q_vec):
        if bf16_primary:
            sharded.search_batch_device(q_vec.view(1, -1), k)
        else:
            sharded.search_device(q_vec, k)

    def step_device(i):
        j = i % n_q
        if args.workload == "blend":
            idx.blend_device(d_q[j:j + 1], d_q2[j:j + 1], d_w[j:j + 1], d_neg[j:j + 1], d_nw[j:j + 1], d_blend)
            search_dev(d_blend[0])
        else:
            search_dev(d_q[j])

    def step_e2e(i):
        return search_via(sharded, host_q[i % n_q])

    sampler = ClockSampler(local_rank) if rank == 0 else None
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
    ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    scan_ms, scans = idx.profile_read()
    kernel_mhz = idx.profile_clock()[0] if bf16_primary else None
    idx.profile(False)
    launches = idx.launch_count - launches0
    reranked = float(idx.batch_stats()[1][0]) if bf16_primary else 0.0

    for i in range(min(args.warmup, 5)):
        step_e2e(i)
    barrier()
    t_wall = time.perf_counter()
    last = None
    for i in range(args.steps):
        last = step_e2e(i)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t_wall) * 1e3) / args.steps
    if sampler:
        sampler.stop()
    assert len(last[0]) == k and np.all(np.diff(last[1]) >= 0)

    scan_avg = max_over_ranks(scan_ms / max(scans, 1))
    row_bytes = BF16_ROW_BYTES if bf16_primary else ROW_BYTES
    bytes_per_query = plan["rows_total"] * row_bytes + world * reranked * ROW_BYTES
    line = None
    if rank == 0:
        peaks = load_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback"
        achieved = n * row_bytes / 1e9 / (scan_avg / 1e3)
        line = {
            "metric": METRIC, "value": bytes_per_query / 1e9 / (ms_step / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 pre-select + f32 re-rank" if bf16_primary else "f32",
            "data": "synthetic" if args.data == "uniform" else "synthetic, clustered (1024 clusters, queries near stored rows)",
            "config": workload_config(args, plan, world),
            "scaling_note": ("value is a RATE (bytes scanned per second), so v_N / (N * v_1) compares per-GPU scan rates; "
                             "the store is BASELINE configs[4] at its stated size (%d rows in total at every N > 1: per-GPU "
                             "rows shrink as N grows), while N = 1 is configs[1] (10M rows: 100M do not fit one GPU); "
                             "fixed-size strong scaling of the 10M-row store is the `strong` sub-record"
                             % plan["rows_total"]),
            "queries_per_s": 1e3 / ms_step,
            "algorithmic_bytes_per_query": bytes_per_query,
            "equivalent_fp32_scan_GBps": plan["rows_total"] * ROW_BYTES / 1e9 / (ms_step / 1e3),
            "e2e": {"value": bytes_per_query / 1e9 / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": ROW_BYTES,
                    "d2h_bytes_per_step": k * 12 + 12, "ms_per_step": e2e_ms, "queries_per_s": 1e3 / e2e_ms,
                    "api": "ShardedIndex.search_batch (nq=1)" if bf16_primary else "ShardedIndex.search"},
            "gpu_launches": int(launches),
            "exchange": exchange_used,
            "store_load_s": load_s,
            "roofline": {"bound": "hbm",
                         "kernel": ("batch_gemm_pair_kernel<FILTER, 64> (bf16 store stream)" if bf16_primary else
                                    ("scan_tma_kernel" if args.variant in (0, 1) else "scan_ldg_kernel")),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8000": achieved / 8000.0, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": n * row_bytes, "avg_launch_ms": scan_avg,
                         "launches_timed": int(scans), "sm_mhz_in_kernel": kernel_mhz,
                         "traffic": None, "traffic_source": None},
            "clocks": sampler.summary() if sampler else None,
        }
        if parity is not None:
            line["parity"] = parity
        if strong is not None:
            line["strong"] = strong

    # ---- configs[2] queries against the sharded store: fused exchange vs NCCL, same process, alternating
    if subs:
        free, _ = torch.cuda.mem_get_info(device)
        need = 0 if bf16_primary else n * BF16_ROW_BYTES + 6_000_000_000
        can = torch.tensor([1 if free > need else 0], device=device)
        dist.all_reduce(can, op=dist.ReduceOp.MIN)
        if int(can[0]):
            rec = measure_sharded_batch(args, torch, dist, idx, sharded, nccl_side, n, device, rank, world, barrier,
                                        max_over_ranks)
            if rec.get("error"):
                return fail("sharded batched search", rec)
            if rank == 0:
                line["configs2_sharded"] = rec
        elif rank == 0:
            line["configs2_sharded"] = {"skipped": "no room for the bf16 copy next to the float32 shard (%d GB free)"
                                                   % (free // 10 ** 9)}
    if rank == 0:
        emit(line)
    idx.close()
    barrier()
    dist.destroy_process_group()
    return 0


def measure_sharded_batch(args, torch, dist, idx, fused_side, nccl_side, rows_per_gpu, device, rank, world, barrier,
                          max_over_ranks):
    """B = 256, k = 100 against the row-sharded store: per rank the tensor-core batched search of its shard, then
    either the peer-memory exchange inside the batched path's last kernel or ONE NCCL all-gather of B*k candidates
    per rank + a merge kernel.  Both are timed in this process, alternating, and must agree bit for bit."""
    B, kb = args.batch, 100
    if not idx.batch_enabled:
        idx.enable_batch()
    rng = np.random.default_rng(99)
    host_q = rng.standard_normal((4, B, DIM), dtype=np.float32)
    host_q /= np.linalg.norm(host_q, axis=2, keepdims=True)
    d_q = torch.from_numpy(host_q).to(device)
    sides = {"fused": fused_side, "nccl": nccl_side} if fused_side.fused else {"nccl": nccl_side}
    # agreement first
    outs = {}
    for name, sh in sides.items():
        ids, d, cnt = sh.search_batch(host_q[0], kb)
        outs[name] = (ids, d.view(np.uint32), cnt)
    agree = len(outs) < 2 or all(np.array_equal(a, b) for a, b in zip(outs["fused"], outs["nccl"]))
    votes = [None] * world
    dist.all_gather_object(votes, bool(agree))
    if not all(votes):
        return {"error": "fused and NCCL batched exchange disagree", "ranks": [r for r, v in enumerate(votes) if not v]}
    steps, warm, reps = max(5, min(args.steps, 10)), 3, 2
    # ONE query through the same pre-selection over the sharded store (what a bf16 copy buys a single search: half
    # the bytes of the float32 scan), checked against the exact sharded scan
    side = fused_side if fused_side.fused else nccl_side
    k1 = args.k
    one_q = d_q[1][:1]
    for _ in range(warm):
        side.search_batch_device(one_q, k1)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        side.search_batch_device(one_q, k1)
    ev1.record()
    barrier()
    one_ms = max_over_ranks(ev0.elapsed_time(ev1)) / steps
    b_ids, b_d, b_n = side.search_batch(host_q[1][:1], k1)
    e_ids, e_d = side.search(host_q[1][0], k1)
    same_one = bool(np.array_equal(b_ids[0, :b_n[0]], e_ids) and
                    np.array_equal(b_d[0, :b_n[0]].view(np.uint32), e_d.view(np.uint32)))
    votes = [None] * world
    dist.all_gather_object(votes, same_one)
    if not all(votes):
        return {"error": "one query through the pre-selection differs from the exact sharded scan"}
    single = {"ms_per_query": one_ms, "queries_per_s": 1e3 / one_ms, "k": k1,
              "scanned_GBps_bf16": rows_per_gpu * world * BF16_ROW_BYTES / 1e9 / (one_ms / 1e3),
              "bit_identical_to_exact_scan": True}
    times = {name: [] for name in sides}
    for _ in range(reps):
        for name, sh in sides.items():
            for i in range(warm):
                sh.search_batch_device(d_q[i % 4], kb)
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for i in range(steps):
                sh.search_batch_device(d_q[i % 4], kb)
            ev1.record()
            barrier()
            times[name].append(max_over_ranks(ev0.elapsed_time(ev1)) / steps)
    best = {name: float(np.mean(v)) for name, v in times.items()}
    winner = min(best, key=best.get)
    return {"single_query_through_bf16_preselect": single, "workload": "B=%d, k=%d batched queries against the %d-row sharded store (configs[2] queries on the configs[4] "
                        "store)" % (B, kb, rows_per_gpu * world),
            "ms_per_step": {name: v for name, v in best.items()}, "ms_per_step_runs": times,
            "queries_per_s": {name: B * 1e3 / v for name, v in best.items()},
            "faster": winner, "steps": steps, "repetitions": reps,
            "fused_equals_nccl_bitwise": len(outs) == 2,
            "default_path": "fused" if fused_side.fused else "nccl"}


def measure_strong_scaling(args, torch, dist, device, rank, local_rank, world, barrier, max_over_ranks, gather_objects):
    """the 10M-row configs[1] store split over the N GPUs (strong scaling): latency per query, the
    in-kernel timeline of the fused exchange (from %globaltimer stamps taken by the scan kernel's last CTA) and the
    NCCL path for comparison.  Every rank generates the same 10M rows and owns a contiguous slice; rank 0 also
    answers from the whole store on one GPU: the sharded answers must equal it bit for bit."""
    from clip_database_b200 import GpuIndex
    from clip_database_b200.sharded import CudaShardBackend, ShardedIndex
    k, total = args.k, 10_000_000
    free = min(gather_objects(torch.cuda.mem_get_info(device)[0]))
    if free < total * ROW_BYTES + 8_000_000_000:
        return {"skipped": "not enough free HBM for the 10M-row store"}
    rows = torch.empty((total, DIM), dtype=torch.float32, device=device)
    generate_store(torch, device, [(1234, total)], 1, into=rows)
    lo, hi = total * rank // world, total * (rank + 1) // world
    idx = GpuIndex(local_rank)
    idx.attach(rows[lo:hi], rowid_base=1 + lo)
    fused = ShardedIndex(CudaShardBackend(idx), fused={"auto": "auto", "fused": True, "nccl": False}[args.exchange])
    nccl = ShardedIndex(CudaShardBackend(idx), fused=False)
    n_q = 32
    host_q = make_queries(rows, n_q, 99, False)
    d_q = torch.from_numpy(host_q).to(device)
    steps, warm = max(20, min(args.steps, 100)), 5
    rec = {"workload": "the 10M-row configs[1] store split over %d GPUs (%d rows per GPU), k=%d" % (world, hi - lo, k),
           "rows_total": total, "rows_per_gpu": hi - lo, "steps": steps, "warmup": warm}

    # single GPU over the whole store (rank 0), for the speed-up and as the answer every rank must reproduce
    whole = None
    if rank == 0:
        one = GpuIndex(local_rank)
        one.attach(rows, rowid_base=1)
        one.use_torch_stream()
        o_ids = torch.empty((1, k), dtype=torch.int64, device=device)
        o_dist = torch.empty((1, k), dtype=torch.float32, device=device)
        o_n = torch.zeros(1, dtype=torch.int32, device=device)
        o_nan = torch.zeros(1, dtype=torch.int64, device=device)
        for i in range(warm):
            one.search_device(d_q[i % n_q].view(1, -1), k, o_ids, o_dist, o_n, o_nan)
        torch.cuda.synchronize()
        rec["one_gpu_ms_per_query"] = timed_loop(torch, 20, lambda i: one.search_device(
            d_q[i % n_q].view(1, -1), k, o_ids, o_dist, o_n, o_nan))
        whole = [one.search(host_q[j], k) for j in range(4)]
        whole = [(r.rowids[0].tolist(), r.distances[0].view(np.uint32).tolist()) for r in whole]
        one.close()
    whole = gather_objects(whole)[0]
    bad = []
    for j in range(4):
        for sh in (fused, nccl):
            ids, d = sh.search(host_q[j], k)
            if ids.tolist() != whole[j][0] or d.view(np.uint32).tolist() != whole[j][1]:
                bad.append({"query": j, "rank": rank, "path": "fused" if sh is fused and fused.fused else "nccl"})
    bad = [b for per_rank in gather_objects(bad) for b in per_rank]
    if bad:
        idx.close()
        return {"error": "sharded answer differs from the single-GPU answer", "where": bad}
    rec["equals_single_gpu_answer_bitwise"] = True

    def timed(sh):
        for i in range(warm):
            sh.search_device(d_q[i % n_q], k)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            sh.search_device(d_q[i % n_q], k)
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1)) / steps

    if fused.fused:
        rec["fused_ms_per_query"] = timed(fused)
        idx.exchange_stats(enable=True, reset=True)
        timed(fused)
        stats, launches = idx.exchange_stats(enable=False, reset=True)
        per_rank = gather_objects(stats)
        rec["timeline_us"] = {
            "what": "mean per launch, from %globaltimer stamps inside the scan kernel: first CTA started -> last CTA "
                    "done with its rows (scan) -> per-SM lists merged (local_merge) -> record stored in every peer's "
                    "inbox (publish) -> every peer's record arrived (wait_peers: includes the skew between the GPUs' "
                    "scans) -> merged and decoded (final_merge)",
            "launches": launches,
            "rank0": per_rank[0],
            "max_over_ranks": {key: max(p[key] for p in per_rank) for key in per_rank[0]},
            "min_over_ranks": {key: min(p[key] for p in per_rank) for key in per_rank[0]}}
        scan_us = float(np.mean([p["scan"] for p in per_rank]))
        rec["scan_GBps_per_gpu_in_kernel"] = (hi - lo) * ROW_BYTES / 1e9 / (scan_us / 1e6) if scan_us > 0 else None
    rec["nccl_ms_per_query"] = timed(nccl)
    best = min(v for key, v in rec.items() if key.endswith("_ms_per_query") and key != "one_gpu_ms_per_query")
    one_ms = gather_objects(rec.get("one_gpu_ms_per_query"))[0]
    rec["one_gpu_ms_per_query"] = one_ms
    rec["speedup_over_one_gpu"] = one_ms / best
    rec["strong_scaling_efficiency"] = one_ms / best / world
    # e2e through the host API
    for i in range(3):
        fused.search(host_q[i], k)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        fused.search(host_q[i % n_q], k)
    barrier()
    rec["e2e_ms_per_query"] = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
    tl = rec.get("timeline_us", {}).get("max_over_ranks")
    if tl:
        tail = tl["local_merge"] + tl["publish"] + tl["wait_peers"] + tl["final_merge"]
        rec["limiter"] = ("scan %.0f us of %.0f us per query; tail (merge + exchange + wait for the slowest peer) %.0f us; "
                          "the rest is launch latency and the gap between launches"
                          % (tl["scan"], rec["fused_ms_per_query"] * 1e3, tail))
    idx.close()
    del rows, fused, nccl
    torch.cuda.empty_cache()
    barrier()
    return rec


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
    if args.workload == "binary":
        return run_binary_workload(args, torch, device, local_rank)
    if world == 1:
        return run_single_gpu(args, torch, device, local_rank)
    dist.init_process_group("nccl", device_id=device)
    if args.workload == "batch":
        return run_sharded_batch_only(args, torch, dist, device, rank, local_rank, world)
    return run_sharded(args, torch, dist, device, rank, local_rank, world)


def run_sharded_batch_only(args, torch, dist, device, rank, local_rank, world):
    """`--workload batch` under torchrun: configs[2]'s queries against a float32 shard per GPU."""
    from clip_database_b200 import GpuIndex
    from clip_database_b200.sharded import CudaShardBackend, ShardedIndex
    n = args.rows or BLOCK_ROWS
    rows = torch.empty((n, DIM), dtype=torch.float32, device=device)
    generate_store(torch, device, [(1234 + rank, n)], 1 + rank * n, into=rows)
    idx = GpuIndex(local_rank)
    idx.attach(rows, rowid_base=1 + rank * n)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])
    fused = ShardedIndex(CudaShardBackend(idx), fused={"auto": "auto", "fused": True, "nccl": False}[args.exchange])
    nccl = ShardedIndex(CudaShardBackend(idx), fused=False)
    rec = measure_sharded_batch(args, torch, dist, idx, fused, nccl, n, device, rank, world, barrier, max_over_ranks)
    if rank == 0:
        best = rec["ms_per_step"][rec["faster"]] if "ms_per_step" in rec else None
        emit({"metric": "knn_batched_queries_per_s", "value": (args.batch * 1e3 / best) if best else None,
              "unit": "queries/s", "n_gpus": world, "higher_is_better": True, "scaling": "weak", **rec})
    idx.close()
    barrier()
    dist.destroy_process_group()
    return 0 if "error" not in rec else 3


if __name__ == "__main__":
    sys.exit(main())
