#!/usr/bin/env python3
"""Load rate of the SQLite -> HBM loaders on one database: the native reader with 1..N connections, and the Python
reader.  Writes a 1M-row reference-schema database to a temp dir first (a child process, ~45 s).
    python tools/loader_bench.py [--rows 1000000]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MAKE = r"""
import sys
sys.path.insert(0, {root!r})
from clip_database_b200 import synth
from oracle import ref
synth.write_reference_db({path!r}, ref.fill_unit_rows({n}, 1152, 1234), binary_codes=False)
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    args = ap.parse_args()
    from clip_database_b200 import GpuIndex, ImageDatabase
    tmp = tempfile.mkdtemp(prefix="clipdb_loader_")
    path = os.path.join(tmp, "big.db")
    subprocess.run([sys.executable, "-c", MAKE.format(root=ROOT, n=args.rows, path=path)], check=True)
    out = {"rows": args.rows, "db_bytes": os.path.getsize(path), "host_cores": os.cpu_count(), "runs": []}
    for readers in (1, 2, 4, 8, 12):
        with GpuIndex(0) as idx:
            idx.set_option("sqlite_readers", readers)
            idx.reserve(args.rows, 1152, explicit_rowids=True)
            t0 = time.perf_counter()
            _, joined = idx.append_sqlite(path)
            dt = time.perf_counter() - t0
            out["runs"].append({"reader": "native", "connections": readers, "rows_per_s": joined / dt,
                                "GBps": joined * 4608 / 1e9 / dt, "seconds": dt})
    t0 = time.perf_counter()
    db = ImageDatabase(path, device=0, native_loader=False)
    out["runs"].append({"reader": "python", "connections": 1, "rows_per_s": args.rows / db.load_seconds,
                        "GBps": args.rows * 4608 / 1e9 / db.load_seconds, "seconds": db.load_seconds})
    db.close()
    os.remove(path)
    os.rmdir(tmp)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
