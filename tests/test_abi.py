"""The C-ABI library loads and exports every symbol include/clipdb.h declares.
No compute calls here (this file runs without a GPU)."""
import ctypes
import os
import re

import pytest

from clip_database_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "clipdb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clipdb_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    names = declared_symbols()
    assert len(names) >= 20
    assert names == sorted(s[0] for s in _lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(build.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), name


def test_abi_version(lib):
    assert lib.clipdb_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path needs a GPU-less host")
    from clip_database_b200 import GpuIndex
    with pytest.raises(_lib.ClipdbError) as e:
        GpuIndex(0)
    assert e.value.code == 2            # CLIPDB_ERR_CUDA, not a silent CPU path


def test_product_does_not_import_the_oracle():
    """Nothing under clip_database_b200/ may reference oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "clip_database_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle_ref" not in src and "vec_shim" not in src, f
