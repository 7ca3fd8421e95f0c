"""Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU).

The built ``libclipdb_b200.so`` sits next to this file so it travels with the
repository snapshot to the GPU box; it is git-ignored (``*.so``).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from typing import List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_NAME = "libclipdb_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)

SOURCES = ["clipdb.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def _inputs() -> List[str]:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(INCLUDE, "clipdb.h"))
    return files


def source_hash() -> str:
    """SHA-256 over the names and contents of everything the library is compiled from.  The same
    string is compiled into the library (``clipdb_source_hash()``), so staleness does not depend on
    file times (which a copy of the tree to another machine does not preserve)."""
    h = hashlib.sha256()
    for f in sorted(_inputs()):
        h.update(os.path.basename(f).encode() + b"\0")
        h.update(open(f, "rb").read())
        h.update(b"\0")
    return h.hexdigest()


def built_hash() -> Optional[str]:
    """The source hash recorded in the built library (read from the file, not by loading it: a stale
    library must not already be mapped when the rebuilt one is dlopen'ed), or None."""
    if not os.path.exists(LIB_PATH):
        return None
    marker = b"clipdb-source-hash:"
    data = open(LIB_PATH, "rb").read()
    at = data.find(marker)
    if at < 0:
        return None
    return data[at + len(marker):at + len(marker) + 64].split(b"\0", 1)[0].decode("ascii", "replace")


def is_stale() -> bool:
    return built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Build ``libclipdb_b200.so`` if missing or compiled from other sources than the ones here.  Safe when
    several processes get here together (one rank per GPU importing the package at the same moment): the build
    is serialised by a file lock, whoever waited re-checks staleness, and the library appears atomically."""
    if not force and not is_stale():
        return LIB_PATH
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():      # another process built it while this one waited
                return LIB_PATH
            tmp = LIB_PATH + ".%d.tmp" % os.getpid()
            cmd = [_nvcc()] + NVCC_FLAGS + ["-DCLIPDB_SOURCE_HASH=\"%s\"" % source_hash(), "-I", INCLUDE, "-o", tmp]
            cmd += [os.path.join(CSRC, s) for s in SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                sys.stderr.write(proc.stdout + proc.stderr)
                raise RuntimeError("nvcc failed building " + LIB_NAME)
            os.replace(tmp, LIB_PATH)
            if verbose:
                sys.stderr.write(proc.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
