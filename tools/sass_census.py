"""Per-kernel SASS opcode census of the shipped library: `cuobjdump -sass libclipdb_b200.so`, counting the
mnemonics that prove the Blackwell paths are really in the binary (B200_PROFILING.md): UTCHMMA / UTCQMMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UTMALDG (TMA tensor loads), UBLKCP
(cp.async.bulk), SYNCS (mbarrier), FFMA2 (packed fp32 fma), plus the plain FFMA / LDG / LDS / POPC counts.
Writes profiles/sass_census.txt.  Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "clip_database_b200", "libclipdb_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FFMA",
         "DFMA", "LDG", "LDS", "STG", "POPC", "SHFL", "ATOM", "RED", "ELECT", "UCGABAR_ARV"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True)
    if out.returncode != 0:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True)
    return out.stdout.splitlines() if out.returncode == 0 else names


def strip_params(name):
    """`void ns::kernel<(bool)0, 256>(Args)` -> `ns::kernel<(bool)0, 256>`: cut at the first '(' outside <>."""
    depth = 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            name = name[:i]
            break
    return name.replace("void ", "").replace("clipdb::", "")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            kernels[cur][op] += 1
            kernels[cur]["__total__"] += 1
    names = demangle(list(kernels))
    lines = ["# SASS opcode census of clip_database_b200/libclipdb_b200.so (tools/sass_census.py; cuobjdump -sass)",
             "# cubins: %s" % ", ".join(arch),
             "# columns: instructions | " + " ".join(WATCH), ""]
    totals = collections.Counter()
    for (mangled, counts), name in zip(kernels.items(), names):
        short = strip_params(name)
        cells = " ".join("%s=%d" % (w, counts[w]) for w in WATCH if counts[w])
        lines.append("%-96s %6d | %s" % (short[:140], counts["__total__"], cells))
        for w in WATCH:
            totals[w] += counts[w]
    lines += ["", "TOTAL " + " ".join("%s=%d" % (w, totals[w]) for w in WATCH if totals[w])]
    out = os.path.join(ROOT, "profiles", "sass_census.txt")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-12:]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
