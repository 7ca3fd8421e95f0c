#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the last `--steps`
repetitions of the per-step kernel sequence, each kernel's time and share of the step.

    python tools/launch_list.py launches.csv --tail 8 [--out profiles/x.txt] [--command "..."]"""
import argparse
import csv
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--tail", type=int, default=8, help="launches in one step (taken from the end of the list)")
    ap.add_argument("--out")
    ap.add_argument("--command", default="")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names = rows[hdr]
    k, g, b = names.index("Kernel Name"), names.index("Grid Size"), names.index("Block Size")
    launches = [r for r in rows[hdr + 1:] if r and r[0].isdigit()]
    step = launches[-args.tail:]
    total = sum(float(r[-1]) for r in step)
    out = ["# ncu launch list (gpu__time_duration.sum, --clock-control none): one step = the last %d launches" % args.tail]
    if args.command:
        out.append("# command: " + args.command)
    out.append("# launches captured: %d" % len(launches))
    for r in step:
        ns = float(r[-1])
        out.append("id %5s %10.2f us %6.2f%%  grid %-14s block %-14s %s" % (r[0], ns / 1e3, 100 * ns / total, r[g], r[b],
                                                                          r[k][:110]))
    out.append("step total %.2f us" % (total / 1e3))
    text = "\n".join(out) + "\n"
    if args.out:
        open(args.out, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
