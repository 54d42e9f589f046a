#!/usr/bin/env python
"""Copy the artefacts of one `tools/gpu_job.sh TAG` run from gpurun_out/ (scratch) into profiles/ (tracked) under round-2
names, rebuild profiles/ncu_figures.json from the raw-page CSVs and write profiles/r02_ncu_summary.md (per-kernel figures
+ per-source-line shares from the source-page CSVs).

    python tools/collect_profiles.py r02j"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))


def last_json_line(path):
    lines = [ln for ln in open(path).read().splitlines() if ln.strip().startswith("{")]
    return json.loads(lines[-1])


def launch_list(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, mi = h.index("Kernel Name"), h.index("Metric Value")
    agg = {}
    for r in rows[hdr + 1:]:
        if len(r) > mi:
            agg.setdefault(r[ki].split("(")[0].replace("void qd::", ""), []).append(float(r[mi].replace(",", "")) / 1e6)
    return {k: v for k, v in agg.items() if k.startswith("qd_")}


def main():
    tag = sys.argv[1]
    sha = open(os.path.join(G, f"sha_{tag}.txt")).read().strip()
    copies = {f"{tag}_bench.json": "r02_bench_n1.json", f"{tag}_bench_reference.json": "r02_bench_reference.json",
              f"{tag}_bench_pathB_4dot.json": "r02_bench_pathB_4dot.json", f"{tag}_sweep.json": "r02_sweep.json",
              f"{tag}_launches_tunnel.csv": "r02_launches_tunnel8.csv", f"{tag}_shell_breakdown.txt": "r02_shell_breakdown.txt",
              f"{tag}_latency.json": "r02_latency.json", f"{tag}_gpu_info.csv": "r02_gpu_info.csv",
              f"{tag}_pytest.txt": "r02_pytest_gpu.txt"}
    for src, dst in copies.items():
        sp = os.path.join(G, src)
        if not os.path.exists(sp):
            print("missing", src)
            continue
        if dst.endswith(".json") and "sweep" not in dst:
            json.dump(last_json_line(sp), open(os.path.join(P, dst), "w"), indent=1)
        else:
            shutil.copy(sp, os.path.join(P, dst))
    lat_err = os.path.join(G, f"{tag}_latency.err")
    if os.path.exists(lat_err):
        with open(os.path.join(P, "r02_latency.json"), "a") as f:
            f.write("\n" + open(lat_err).read().strip().splitlines()[-1] + "\n")
    figs = os.path.join(P, "ncu_figures.json")
    if os.path.exists(figs):
        os.remove(figs)
    caps = [("fast8", "qd_scan_fast", 58720256, "qd_scan_kernel<8,default>"),
            ("select8", "qd_tunnel_select2", 1835008, "qd_tunnel_select2_kernel<8>"),
            ("eigen8", "qd_tunnel_eigen2", 1835008, "qd_tunnel_eigen2_kernel<8>"),
            ("select4", "qd_tunnel_select2", 3145728, "qd_tunnel_select2_kernel<4>"),
            ("eigen4", "qd_tunnel_eigen2", 3145728, "qd_tunnel_eigen2_kernel<4>"),
            ("hh8", "qd_tunnel_eigen_kernel", 1835008, "qd_tunnel_eigen_kernel<8>"),
            ("blk8", "qd_tunnel_select_kernel", 1835008, "qd_tunnel_select_kernel<8>")]
    md = [f"# ncu summary r02 (source hash {sha}; captures of `tools/gpu_job.sh {tag}`)\n",
          "`ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1`, each after the same command exited 0 "
          "without ncu; raw and source pages exported on the box (`tools/ncu_export.sh`), the 25 MB reports stay there.\n"]
    for name, match, pixels, key in caps:
        raw = os.path.join(G, f"prof_{tag}_{name}_raw.csv")
        if not os.path.exists(raw):
            continue
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_figures.py"), raw, "--kernel", key, "--match", match,
                              "--pixels", str(pixels), "--sha", sha], capture_output=True, text=True)
        fig = json.load(open(figs))[key]
        md.append(f"\n## {key}  ({pixels / 1e6:.2f} Mpixel per launch)\n\n```")
        for k in ("warp_instr_per_pixel", "dram_bytes_per_pixel", "issue_active_pct", "warps_active_pct", "registers",
                  "ms_per_launch_under_ncu", "pipe_fp64_pct", "pipe_alu_pct", "pipe_lsu_pct"):
            md.append(f"{k:28s} {fig[k]:.2f}" if isinstance(fig[k], float) else f"{k:28s} {fig[k]}")
        md.append("```")
        src = os.path.join(G, f"prof_{tag}_{name}_src.csv.gz")
        if os.path.exists(src):
            fname = ("qd_kernels.cuh" if name == "fast8" else "qd_tunnel_noda.cuh" if name.startswith("eigen") else
                     "qd_tunnel_enum.cuh" if name.startswith("select") else "qd_tunnel.cuh")
            top = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), src, "--file", fname, "--top", "14"],
                                 capture_output=True, text=True).stdout
            md.append("\n```\n" + top.strip() + "\n```")
    ll = os.path.join(G, f"{tag}_launches_tunnel.csv")
    if os.path.exists(ll):
        md.append("\n## Tunnel pipeline, per-kernel durations (ms, last launch of each; `ncu --metrics gpu__time_duration.sum`, "
                  "8 dots, 1.84 Mpixel)\n\n```\n" + json.dumps(launch_list(ll), indent=1) + "\n```")
    open(os.path.join(P, "r02_ncu_summary.md"), "w").write("\n".join(md) + "\n")
    print(open(figs).read()[:600])


if __name__ == "__main__":
    main()
