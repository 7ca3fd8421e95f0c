"""examples/abi_smoke.c: the C ABI used from plain C (no Python, no torch) — what a C / cgo / JNI
binding of the reference would link against."""
import os
import subprocess

import pytest

from clip_database_b200 import build as cuda_build
from conftest import have_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compile_example(tmp_path) -> str:
    cuda_build.build()
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(cuda_build.LIB_PATH)
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lclipdb_b200",
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
