#!/usr/bin/env python
"""Iteration statistics of qd_tunnel_eigen2_kernel from a -DQD_E2_STATS build (the kernel then writes its counters in
place of <n>): QDSIM_LIB=variants/libqdsim_stats.so python tools/e2_stats.py --cases 4:64,8:32"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rl-agent-for-qubit-array-tuning_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from qdsim import N_F64, Engine, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="4:64,8:32")
    ap.add_argument("--res", type=int, default=64)
    a = ap.parse_args()
    os.environ["QDSIM_EIGEN"] = "noda"
    eng = Engine(0)
    for case in a.cases.split(","):
        n_dot, n_env = map(int, case.split(":"))
        dev = synth.sample_barrier_devices(n_env, n_dot, seed=1234)
        mb = synth.tunnel_batch(dev, latching=False, noise=False)
        eng.set_models(mb)
        scans = synth.env_step_scans(mb, dev, res=a.res, seed=99, radial=False)
        pixels = len(scans) * a.res * a.res
        z = torch.empty(pixels, dtype=torch.float32, device="cuda")
        n = torch.empty((pixels, n_dot), dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream()
        eng.scan_upload(scans, st)
        eng.scan_launch(z, n, N_F64, 0, st)
        torch.cuda.synchronize()
        s = n.cpu().numpy()
        fb = np.isnan(s[:, 0])
        s = s[~fb]
        hist = collections.Counter(zip(s[:, 0].astype(int).tolist(), s[:, 1].astype(int).tolist()))
        top = sorted(hist.items(), key=lambda kv: -kv[1])[:14]
        col = np.arange(pixels)[~fb] % a.res
        print(json.dumps({"n_dot": n_dot, "pixels": pixels, "fallback": float(fb.mean()), "fac": float(s[:, 0].mean()),
                          "iters": float(s[:, 1].mean()), "aggfail": float(s[:, 2].mean()), "testfail": float(s[:, 3].mean()),
                          "fac_first_col": float(s[col == 0, 0].mean()), "fac_other_cols": float(s[col != 0, 0].mean()),
                          "hist(fac,it)": [[list(k), v / len(s)] for k, v in top]}), flush=True)


if __name__ == "__main__":
    main()
