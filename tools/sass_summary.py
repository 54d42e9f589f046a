#!/usr/bin/env python
"""Static evidence of what was built: per hot kernel of libqdsim.so, registers / stack / shared memory (cuobjdump
-res-usage) and SASS mnemonic counts (cuobjdump -sass): UBLKCP (1-D TMA bulk copy), SYNCS (mbarrier), DFMA / DADD / DMUL,
MUFU, SHFL / VOTE / MATCH / REDUX, local-memory LDL / STL (spills), and -- for the record -- the absence of tensor-core
and tensor-TMA instructions (UTMALDG, UTCHMMA, HMMA): the path has no contraction shape.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rl-agent-for-qubit-array-tuning_b200", "csrc", "libqdsim.so")
HOT = ["qd_scan_fast_kernelILi8E", "qd_scan_fast_kernelILi4E", "qd_scan_kernelILi8ELi0ELb0ELb0E", "qd_scan_kernelILi6ELi2ELb0ELb0E",
       "qd_scan_kernelILi8ELi3ELb0ELb0E", "qd_tunnel_relax_kernelILi8E", "qd_tunnel_select_kernelILi8E", "qd_tunnel_select2_kernelILi8E", "qd_tunnel_select2_kernelILi4E", "qd_tunnel_eigen2_kernelILi8E", "qd_tunnel_eigen_kernelILi8E",
       "qd_tunnel_select_kernelILi4E", "qd_tunnel_eigen2_kernelILi4E", "qd_tunnel_eigen_kernelILi4E", "qd_tunnel_gs_kernelILi8E", "qd_normalise_reg_kernelIhE",
       "qd_normalise_reg_kernelIfE", "qd_build_q_kernel"]
WATCH = ["UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "DSETP", "MUFU", "SHFL", "VOTE", "MATCH", "REDUX", "LDS", "STS", "LDG", "STG", "ATOMS",
         "LDL", "STL", "UTMALDG", "UTCHMMA", "UTCQMMA", "HMMA", "LDTM"]


def main():
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    print("# SASS summary of", os.path.relpath(LIB, ROOT))
    print("# embedded ELF images:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))) or "?")
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in ln:
            usage[cur] = ln.strip()
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    print(f"# {len(usage)} kernels in the library; hot ones below\n")
    for blk in blocks[1:]:
        name = blk.split("\n", 1)[0].strip()
        if not any(h in name for h in HOT):
            continue
        ops = Counter()
        total = 0
        for ln in blk.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
            if m:
                total += 1
                ops[m.group(1)] += 1
        print(name)
        print("   ", usage.get(name, ""))
        print("    instructions:", total, " ".join(f"{k}={ops[k]}" for k in WATCH if ops[k] or k in ("UBLKCP", "LDL", "STL", "UTMALDG", "UTCHMMA", "HMMA")))
    all_ops = Counter(re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", sass))
    print("\n# whole library:", " ".join(f"{k}={all_ops[k]}" for k in ("UBLKCP", "SYNCS", "DFMA", "UTMALDG", "UTCHMMA", "UTCQMMA", "HMMA", "LDTM")))


if __name__ == "__main__":
    main()
