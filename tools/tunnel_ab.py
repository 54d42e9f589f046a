#!/usr/bin/env python
"""A/B of the two eigen stages of the tunnel path on the same batch: QDSIM_EIGEN=householder (tridiagonalisation +
multisection) against the default (Noda iteration + fix-up); prints the differences in <n> and both timings.

    python tools/tunnel_ab.py [--cases 4:256,8:64] [--steps 3]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rl-agent-for-qubit-array-tuning_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from qdsim import N_F64, Engine, synth  # noqa: E402


VAR = "QDSIM_EIGEN"


def run(eng, scans, pixels, n_dot, steps, mode):
    if mode:
        os.environ[VAR] = mode
    else:
        os.environ.pop(VAR, None)
    z = torch.empty(pixels, dtype=torch.float32, device="cuda")
    n = torch.empty((pixels, n_dot), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    eng.scan_upload(scans, st)
    for _ in range(2):
        eng.scan_launch(z, n, N_F64, 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        eng.scan_launch(z, n, N_F64, 0, st)
    e1.record(st)
    torch.cuda.synchronize()
    return n.cpu().numpy(), e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="4:256,8:64")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--var", default="QDSIM_EIGEN", help="environment switch to A/B (QDSIM_EIGEN or QDSIM_SELECT)")
    ap.add_argument("--ref", default="householder", help="value of the switch for the reference run (QDSIM_SELECT: block)")
    a = ap.parse_args()
    global VAR
    VAR = a.var
    eng = Engine(0)
    for case in a.cases.split(","):
        n_dot, n_env = map(int, case.split(":"))
        dev = synth.sample_barrier_devices(n_env, n_dot, seed=1234)
        mb = synth.tunnel_batch(dev, latching=False, noise=False)
        eng.set_models(mb)
        scans = synth.env_step_scans(mb, dev, res=a.res, seed=99, radial=False)
        pixels = len(scans) * a.res * a.res
        n_h, ms_h = run(eng, scans, pixels, n_dot, a.steps, a.ref)
        n_n, ms_n = run(eng, scans, pixels, n_dot, a.steps, "")
        d = np.abs(n_h - n_n).max(axis=1)
        bad = np.nonzero(~(d < 1e-6))[0]
        print(json.dumps({"n_dot": n_dot, "n_env": n_env, "pixels": pixels, "ms_householder": ms_h, "ms_noda": ms_n,
                          "Mpix_s_householder": pixels / ms_h / 1e3, "Mpix_s_noda": pixels / ms_n / 1e3,
                          "nan": int(np.isnan(n_n).any(axis=1).sum()), "max_abs": float(np.nanmax(d)),
                          "frac_gt_1e-9": float((d > 1e-9).mean()), "frac_gt_1e-7": float((d > 1e-7).mean()),
                          "n_gt_1e-6": int(len(bad)), "first_bad": bad[:8].tolist(),
                          "bad_vals": [[n_h[i].tolist(), n_n[i].tolist()] for i in bad[:3]]}), flush=True)


if __name__ == "__main__":
    main()
