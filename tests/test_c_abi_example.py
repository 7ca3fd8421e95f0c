"""examples/abi_smoke.c: the C ABI used from plain C (no Python, no torch) — what a C / cgo / JNI
binding of the reference would link against."""
import os
import subprocess

import pytest

from clip_database_b200 import build as cuda_build
from conftest import have_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compile_example(tmp_path, name: str = "abi_smoke") -> str:
    cuda_build.build()
    exe = str(tmp_path / name)
    libdir = os.path.dirname(cuda_build.LIB_PATH)
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", name + ".c"), "-o", exe, "-L", libdir, "-lclipdb_b200",
           "-Wl,-rpath," + libdir, "-lm"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return exe


def test_c_program_links_and_fails_loudly_without_a_device(tmp_path):
    if have_gpu():
        pytest.skip("a CUDA device is visible: covered by the gpu test")
    proc = subprocess.run([compile_example(tmp_path)], capture_output=True, text=True, timeout=120)
    assert proc.returncode == 2 and "no CPU fallback" in proc.stderr


@pytest.mark.gpu
def test_c_program_matches_its_own_brute_force(tmp_path):
    assert have_gpu()
    proc = subprocess.run([compile_example(tmp_path)], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "abi_smoke ok" in proc.stdout


def test_sqlite_example_compiles_and_needs_a_device(tmp_path):
    exe = compile_example(tmp_path, "abi_sqlite")
    if have_gpu():
        pytest.skip("a CUDA device is visible: covered by the gpu test")
    q = tmp_path / "q.f32"
    q.write_bytes(b"\0" * 4608)
    proc = subprocess.run([exe, str(tmp_path / "none.db"), str(q)], capture_output=True, text=True, timeout=120)
    assert proc.returncode == 2 and "no CPU fallback" in proc.stderr


@pytest.mark.gpu
def test_sqlite_example_answers_like_the_reference_statement(tmp_path):
    """examples/abi_sqlite.c: SQLite file -> native reader -> search, from plain C, against the reference's own
    statement executed by the real SQLite on the same file."""
    assert have_gpu()
    import numpy as np
    from clip_database_b200 import synth
    from oracle import sql_harness
    exe = compile_example(tmp_path, "abi_sqlite")
    rows = synth.unit_rows(9000, 1152, 5)
    db = str(tmp_path / "c.db")
    synth.write_reference_db(db, rows, drop_mapping_for=[4])
    query = synth.unit_rows(1, 1152, 6)[0]
    (tmp_path / "q.f32").write_bytes(query.astype("<f4").tobytes())
    proc = subprocess.run([exe, db, str(tmp_path / "q.f32"), "12"], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "loaded 8999 of 9000 vec0 rows" in proc.stderr
    got = [line.split("\t") for line in proc.stdout.strip().splitlines()]
    want = sql_harness.reference_search(db, query, 12)
    assert [p for _, p in got] == [p for p, _ in want]
    assert np.allclose([float(s) for s, _ in got], [s for _, s in want], atol=2e-6)
