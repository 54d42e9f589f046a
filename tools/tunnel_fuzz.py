#!/usr/bin/env python
"""Random batches through both forms of the tunnel path's select and eigen stages (new kernels vs QDSIM_SELECT=block +
QDSIM_EIGEN=householder): sizes, dot counts, offsets, capacitance models; prints one line per case and fails on a NaN or on
more than 1e-4 of the pixels differing by more than 1e-7.

    python tools/tunnel_fuzz.py [--cases 12] [--seed 0]"""
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


def run(eng, scans, pixels, n_dot, new):
    for var, val in (("QDSIM_SELECT", "block"), ("QDSIM_EIGEN", "householder")):
        if new:
            os.environ.pop(var, None)
        else:
            os.environ[var] = val
    z = torch.empty(pixels, dtype=torch.float32, device="cuda")
    n = torch.empty((pixels, n_dot), dtype=torch.float64, device="cuda")
    eng.scan_open(scans, z, n, N_F64, 0)
    torch.cuda.synchronize()
    return n.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=12)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    eng = Engine(0)
    bad = 0
    for c in range(a.cases):
        n_dot = int(rng.integers(4, 9))
        res = int(rng.choice([17, 24, 33, 48, 64]))
        n_env = int(rng.integers(4, 40))
        off = float(rng.choice([1.0, 2.5, 5.0, 8.0]))
        dev = synth.sample_barrier_devices(n_env, n_dot, seed=int(rng.integers(1 << 30)))
        mb = synth.tunnel_batch(dev, latching=False, noise=False)
        if rng.random() < 0.3:
            mb.params["vc_alpha"] = 0.02
            mb.params["vc_beta"] = 0.01
        eng.set_models(mb)
        scans = synth.env_step_scans(mb, dev, res=res, seed=int(rng.integers(1 << 30)), offset_range=off, radial=False)
        pixels = len(scans) * res * res
        n_new = run(eng, scans, pixels, n_dot, True)
        n_old = run(eng, scans, pixels, n_dot, False)
        d = np.abs(n_new - n_old).max(axis=1)
        frac = float((d > 1e-7).mean())
        ok = bool(np.isfinite(n_new).all() and frac < 1e-4)
        bad += not ok
        print(json.dumps({"case": c, "n_dot": n_dot, "res": res, "n_env": n_env, "offset": off, "pixels": pixels,
                          "nan": int((~np.isfinite(n_new)).sum()), "max_abs": float(np.nanmax(d)), "frac_gt_1e-7": frac,
                          "median": float(np.median(d)), "ok": ok}), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
