#!/usr/bin/env python
"""Turn an `ncu --set full` capture of one kernel into the per-pixel figures bench.py reports (profiles/ncu_figures.json):

    python tools/ncu_figures.py REP --kernel "qd_scan_kernel<8,default>" --pixels 58720256 --sha $(cat gpurun_out/sha.txt)

warp instructions per pixel (smsp__inst_executed.sum), DRAM bytes per pixel (dram__bytes_read.sum + dram__bytes_write.sum),
issue-active %, registers, duration -- each keyed to the hash of the kernel sources it was captured from (bench.source_sha),
so bench.py can say when a figure is stale."""
import argparse
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"smsp__inst_executed.sum": "inst", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue", "gpu__time_duration.sum": "dur",
        "launch__registers_per_thread": "regs", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def raw(rep):
    if rep.endswith(".csv"):                         # a raw-page CSV exported on the GPU box
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(head, units, r):
            if h in WANT:
                try:
                    d[WANT[h]] = float(v.replace(",", "")) * SCALE.get(u, 1.0)
                except ValueError:
                    pass
            elif h == "Kernel Name":
                d["name"] = v
        recs.append(d)
    return recs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", required=True, help="key in ncu_figures.json, e.g. 'qd_scan_kernel<8,default>'")
    ap.add_argument("--match", default=None, help="substring of the ncu kernel name (default: text before '<')")
    ap.add_argument("--pixels", type=float, required=True, help="pixels one captured launch processes")
    ap.add_argument("--sha", required=True, help="bench.source_sha() of the profiled tree")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ncu_figures.json"))
    a = ap.parse_args()
    match = a.match or a.kernel.split("<")[0]
    recs = [r for r in raw(a.rep) if match in r.get("name", "")]
    if not recs:
        raise SystemExit(f"no launch of {match} in {a.rep}")
    r = recs[-1]
    fig = {"warp_instr_per_pixel": r["inst"] / a.pixels, "dram_bytes_per_pixel": (r["rd"] + r["wr"]) / a.pixels,
           "issue_active_pct": r.get("issue"), "registers": r.get("regs"), "ms_per_launch_under_ncu": r.get("dur"),
           "warps_active_pct": r.get("warps"), "pipe_fp64_pct": r.get("fp64"), "pipe_alu_pct": r.get("alu"),
           "pipe_lsu_pct": r.get("lsu"), "pixels_per_launch": a.pixels, "source_sha": a.sha,
           "source": f"ncu --set full --clock-control none, {os.path.basename(a.rep)}"}
    try:
        allfig = json.load(open(a.out))
    except (OSError, ValueError):
        allfig = {}
    allfig[a.kernel] = fig
    json.dump(allfig, open(a.out, "w"), indent=1, sort_keys=True)
    print(json.dumps({a.kernel: fig}, indent=1))


if __name__ == "__main__":
    main()
