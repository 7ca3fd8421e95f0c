#!/usr/bin/env python3
"""Turn `ncu -i X.ncu-rep --page raw --csv` output into the small tables kept under profiles/.

    python tools/ncu_summary.py raw.csv [--md out.md] [--csv out_raw.csv] [--title "..."] [--command "..."]

Writes the selected metrics (one column per captured launch) as a markdown table, and a trimmed
CSV with the same metrics so the numbers quoted in DESIGN.md can be re-read without the report."""
import argparse
import csv
import sys

WANT = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic",
    "gpc__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("--md")
    ap.add_argument("--csv")
    ap.add_argument("--title", default="ncu --set full")
    ap.add_argument("--command", default="")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    launches = [r for r in rows[hdr + 2:] if r and r[0].isdigit()]
    cols = []
    for want in WANT:
        hit = [i for i, n in enumerate(names) if n == want or n.endswith("." + want)]
        if hit:
            cols.append((want, hit[0]))
    kcol = names.index("Kernel Name")
    out = ["# " + args.title, ""]
    if args.command:
        out += ["command: `%s`" % args.command, ""]
    for j, r in enumerate(launches):
        out.append("launch %d: `%s` grid %s block %s" % (j + 1, r[kcol][:160], r[names.index("Grid Size")],
                                                          r[names.index("Block Size")]))
    out += ["", "| metric | unit | " + " | ".join("launch %d" % (j + 1) for j in range(len(launches))) + " |",
            "|---|---|" + "---|" * len(launches)]
    for want, i in cols:
        out.append("| %s | %s | %s |" % (want, units[i], " | ".join(r[i] for r in launches)))
    text = "\n".join(out) + "\n"
    if args.md:
        open(args.md, "w").write(text)
    else:
        sys.stdout.write(text)
    if args.csv:
        with open(args.csv, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit"] + ["launch %d" % (j + 1) for j in range(len(launches))])
            for want, i in cols:
                w.writerow([want, units[i]] + [r[i] for r in launches])


if __name__ == "__main__":
    main()
