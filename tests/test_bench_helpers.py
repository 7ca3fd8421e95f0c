"""bench.py's host-side logic that decides WHAT is measured: the store plan per GPU count, the generation
blocks, and the comparison against the float64 ground truth."""
import argparse
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(bench)


def args(**kw):
    base = dict(rows=0, store="auto", k=20, workload="single")
    base.update(kw)
    return argparse.Namespace(**base)


def with_box(monkeypatch, hbm_gb, host_gb):
    monkeypatch.setattr(bench, "gpu_total_bytes", lambda index=0: int(hbm_gb * 1e9))
    monkeypatch.setattr(bench, "host_mem_available_bytes", lambda: int(host_gb * 1e9))


def test_plan_holds_100m_rows_at_every_gpu_count(monkeypatch):
    with_box(monkeypatch, 191.5, 2000)
    for n, kind in ((8, "fp32"), (4, "fp32"), (2, "bf16-primary")):
        plan = bench.plan_store(args(), n)
        assert plan["rows_total"] == 100_000_000 and plan["rows_per_gpu"] == 100_000_000 // n
        assert plan["store"] == kind and "note" not in plan
        assert plan["hbm_fp32_rows"] + plan["host_fp32_rows"] == plan["rows_per_gpu"]
    two = bench.plan_store(args(), 2)
    assert two["hbm_fp32_rows"] % 128 == 0 and two["host_fp32_rows"] > 0
    # what stays in HBM fits it: bf16 copy + float32 tier + the reserve
    assert two["rows_per_gpu"] * 2304 + two["hbm_fp32_rows"] * 4608 <= 191.5e9 - 14e9
    assert bench.plan_store(args(), 1)["rows_total"] == 10_000_000


def test_plan_shrinks_to_the_host_memory_and_says_so(monkeypatch):
    with_box(monkeypatch, 191.5, 251)
    plan = bench.plan_store(args(), 2)
    assert plan["store"] == "bf16-primary" and plan["rows_per_gpu"] < 50_000_000 and "note" in plan
    assert plan["rows_per_gpu"] % bench.CHUNK_ROWS == 0
    assert plan["host_fp32_rows"] * 4608 * 2 <= 0.8 * 251e9
    cfg = bench.workload_config(args(), plan, 2)
    assert cfg["rows_total"] == 2 * plan["rows_per_gpu"] and cfg["scanned_bytes_per_row"] == 2304 and "note" in cfg


def test_both_arms_print_the_same_config(monkeypatch):
    with_box(monkeypatch, 191.5, 2000)
    for n in (1, 2, 4, 8):
        a = bench.workload_config(args(), bench.plan_store(args(), n), n)
        b = bench.workload_config(args(), bench.plan_store(args(), n), n)
        assert a == b and a["k"] == 20


def test_blocks_are_the_same_data_for_every_gpu_count(monkeypatch):
    with_box(monkeypatch, 191.5, 2000)
    seen = {}
    for n in (2, 4, 8):
        plan = bench.plan_store(args(), n)
        seeds = []
        for rank in range(n):
            blocks = bench.rank_blocks(args(), plan, rank, n)
            assert sum(r for _, r in blocks) == plan["rows_per_gpu"]
            seeds += [s for s, _ in blocks]
        seen[n] = seeds
    assert seen[2] == seen[4] == seen[8] == [1234 + b for b in range(8)]
    assert bench.rank_blocks(args(rows=1000), {"rows_per_gpu": 1000}, 3, 4) == [(1237, 1000)]


def test_comparison_with_the_ground_truth():
    truth_ids = np.arange(1, 41)
    truth_d = np.linspace(0.5, 0.9, 40)
    truth_d[5] = truth_d[4] + 1e-9                         # a near-tie between ranks 4 and 5
    got_ids = truth_ids[:20].copy()
    got_d = truth_d[:20].astype(np.float32)
    assert bench.compare_with_truth(got_ids, got_d, truth_d, truth_ids, 20)[:2] == (True, 0)
    swapped = got_ids.copy()
    swapped[[4, 5]] = swapped[[5, 4]]                      # a tie within the tolerance may swap
    same, beyond, max_abs, ok = bench.compare_with_truth(swapped, got_d, truth_d, truth_ids, 20)
    assert (same, beyond, ok) == (False, 0, True) and max_abs < 1e-6
    wrong = got_ids.copy()
    wrong[0] = 30                                          # a row that does not belong there
    assert bench.compare_with_truth(wrong, got_d, truth_d, truth_ids, 20)[1] >= 1
    stranger = got_ids.copy()
    stranger[3] = 999                                      # not even among the best k + slack
    same, beyond, _, ok = bench.compare_with_truth(stranger, got_d, truth_d, truth_ids, 20)
    assert beyond >= 1 and not ok
    off = got_d.copy()
    off[7] += 1e-3
    assert bench.compare_with_truth(got_ids, off, truth_d, truth_ids, 20)[3] is False


def test_parity_queries_sit_next_to_the_plant():
    q = bench.parity_queries(8)
    assert q.shape == (8, 1152) and np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-6)
    assert float(q[0] @ bench.plant_vector()) > 0.99
    assert abs(float(q[1] @ bench.plant_vector())) < 0.2


def test_float64_truth_streams_like_a_full_sort():
    """The running (distance, rowid) top-k over chunks == one stable sort over everything, ties included."""
    import torch
    rng = np.random.default_rng(3)
    rows = rng.standard_normal((6000, 1152)).astype(np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    rows[4000] = rows[17]
    rows[5999] = rows[17]                                  # exact ties across chunks
    queries = np.stack([rows[17], rows[100] + 0.1 * rows[200]]).astype(np.float32)
    k = 20
    truth = bench.Float64Truth(torch, torch.device("cpu"), queries, k)
    for lo in range(0, 6000, 2500):                        # uneven chunks
        truth.update(torch.from_numpy(rows[lo:lo + 2500]), 1 + lo)
    d, ids = truth.host()
    r64, q64 = rows.astype(np.float64), queries.astype(np.float64)
    for j in range(2):
        full = 1.0 - (r64 @ q64[j]) / (np.linalg.norm(r64, axis=1) * np.linalg.norm(q64[j]))
        order = np.lexsort((np.arange(6000), full))[:k + 16]
        assert ids[j].tolist() == (order + 1).tolist()
        assert np.allclose(d[j], full[order], atol=1e-12)
    assert ids[0][:3].tolist() == [18, 4001, 6000]
    # merging two shards' lists == one list over both
    a = bench.Float64Truth(torch, torch.device("cpu"), queries, k)
    b = bench.Float64Truth(torch, torch.device("cpu"), queries, k)
    a.update(torch.from_numpy(rows[:3000]), 1)
    b.update(torch.from_numpy(rows[3000:]), 3001)
    a.merge_from([b.host()])
    assert np.array_equal(a.host()[1], ids)
