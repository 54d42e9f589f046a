#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals from an ncu report captured with --import-source on:

    python tools/ncu_lines.py gpurun_out/X.ncu-rep [--top 40] [--ranges "name:lo-hi,..."] [--file qd_tunnel.cuh]

Reads `ncu -i REP --page source --print-source cuda,sass --csv` and sums the SASS rows under each CUDA line."""
import argparse
import csv
import subprocess
from collections import defaultdict


def load(rep):
    if rep.endswith(".gz"):                          # a source-page CSV exported on the GPU box (the .ncu-rep stays there)
        import gzip
        out = gzip.open(rep, "rt").read()
    elif rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                             capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    per = defaultdict(lambda: [0, 0, ""])           # (file, line) -> [inst, samples, text]
    cur_file, cur_line = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] in ("Line No", "Function Name", "Kernel Name"):
            continue
        if r[0] != "":                               # a CUDA line; its text may contain commas -> variable columns
            cur_line = int(r[0])
            per[(cur_file, cur_line)][2] = ",".join(r[1:]).split(",-,-,")[0].strip()
            continue
        try:                                         # SASS row: ['', '', addr, sass, all, not-issued, samples, inst, ...]
            inst = int(r[7])
            samples = int(r[6])
        except (ValueError, IndexError):
            continue
        per[(cur_file, cur_line)][0] += inst
        per[(cur_file, cur_line)][1] += samples
    return per


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--file", default=None)
    ap.add_argument("--ranges", default=None)
    a = ap.parse_args()
    per = load(a.rep)
    ti = sum(v[0] for v in per.values())
    ts = sum(v[1] for v in per.values())
    print(f"total inst {ti}  samples {ts}")
    byfile = defaultdict(lambda: [0, 0])
    for (f, l), v in per.items():
        byfile[f][0] += v[0]
        byfile[f][1] += v[1]
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:24s} inst {100 * v[0] / ti:6.2f}%  samples {100 * v[1] / max(ts, 1):6.2f}%")
    if a.ranges:
        for spec in a.ranges.split(","):
            name, rng = spec.split(":")
            lo, hi = map(int, rng.split("-"))
            i = sum(v[0] for (f, l), v in per.items() if f == a.file and lo <= l <= hi)
            s = sum(v[1] for (f, l), v in per.items() if f == a.file and lo <= l <= hi)
            print(f"  {name:24s} {lo:4d}-{hi:4d}  inst {100 * i / ti:6.2f}%  samples {100 * s / max(ts, 1):6.2f}%")
    print("--- top lines")
    items = [(k, v) for k, v in per.items() if a.file is None or k[0] == a.file]
    for (f, l), v in sorted(items, key=lambda kv: -kv[1][0])[:a.top]:
        print(f"{100 * v[0] / ti:6.2f}% inst {100 * v[1] / max(ts, 1):6.2f}% smp  {f}:{l}  {v[2][:110]}")


if __name__ == "__main__":
    main()
