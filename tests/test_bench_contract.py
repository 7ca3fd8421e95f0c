"""bench.py's reference arm runs on the host CPU, so its JSON contract can be checked here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                           "--warmup", "1", "--cpu-sample-rows", "12000", *extra],
                          capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    return proc.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "knn_scan_throughput" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["config"]["workload"].startswith("single-query cosine KNN")


def test_reference_arm_under_torchrun_only_rank0_prints():
    assert run("--gpus", "4", env={"RANK": "2", "WORLD_SIZE": "4", "LOCAL_RANK": "2"}).strip() == ""
    d = json.loads(run("--gpus", "4", env={"RANK": "0", "WORLD_SIZE": "4", "LOCAL_RANK": "0"}))
    assert d["n_gpus"] == 4 and d["config"]["rows_per_gpu"] == 25_000_000 and d["config"]["rows_total"] == 100_000_000
