"""Generate the golden vectors under tests/golden/ with the CPU oracle ("restated-reference" goldens: the reference
package qarray==1.6.0 is not installable here).  Run from the repo root:

    python tests/golden/make_golden.py

Each .npz holds the raw inputs (non-Maxwell capacitances, per-env parameters, qd_scan records as bytes) and the oracle
outputs (z float64, n float64, margin = gap between the two lowest candidate energies)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "rl-agent-for-qubit-array-tuning_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from qdsim import (FLAG_CARRY_ROWS, FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL, FLAG_THERMAL, synth)  # noqa: E402
from util import oracle_batch  # noqa: E402

CASES = {
    # name: (n_dot, n_env, res, algorithm, flags, kwargs)
    "c1_2dot_64x64_noise_free": (2, 1, 64, "default", 0, dict(latching=False, noise=False)),
    "c2_4dot_latched_full_noise": (4, 2, 32, "default", FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL, dict()),
    "c2b_4dot_flat_pass": (4, 1, 32, "default", FLAG_LATCH | FLAG_NOISE | FLAG_CARRY_ROWS, dict()),
    "c3_6dot_brute_force": (6, 1, 12, "brute_force", 0, dict(latching=False, noise=False)),
    "c4_8dot_latched_full_noise": (8, 1, 32, "default", FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL, dict()),
    "t_3dot_thermal": (3, 2, 32, "default", FLAG_THERMAL, dict(latching=False, noise=False, thermal=True)),
    "t_5dot_thresholded": (5, 1, 32, "thresholded", FLAG_LATCH, dict(noise=False, threshold=0.6)),
    # Path B (tunnel-coupled, what env.step runs in barrier mode): raw barrier matrices are stored too
    "b_4dot_tunnel_latched_noise": (4, 1, 24, "tunnel", FLAG_LATCH | FLAG_NOISE, dict()),
    "b_6dot_tunnel_coupled": (6, 1, 10, "tunnel", 0, dict(latching=False, noise=False)),
}


def main():
    for name, (n_dot, n_env, res, alg, flags, kw) in CASES.items():
        seed = 1000 + sum(map(ord, name))
        if alg == "tunnel":
            dev = synth.sample_barrier_devices(n_env, n_dot, seed=seed)
            mb = synth.tunnel_batch(dev, **kw)
        else:
            dev = synth.sample_devices(n_env, n_dot, seed=seed)
            mb = synth.model_batch(dev, algorithm=alg, **kw)
        if flags & FLAG_NOISE:
            mb.params["tele_p01"] = 0.03
            mb.params["tele_p10"] = 0.08
            mb.params["tele_amp"] = 0.01
        scans = synth.env_step_scans(mb, dev, res=res, seed=seed + 1, offset_range=3.0)
        if alg == "tunnel":
            scans = scans[:2].copy()
        if n_dot == 2:                      # BASELINE config 1: a single do2d_open window
            scans = scans[:1]
        if flags & FLAG_RADIAL:
            scans["rad_zero_radius"] = 1.0
            scans["rad_alpha"] = 0.02
        z, n, margin = oracle_batch(mb, scans, flags)
        extra = {k: dev[k] for k in ("Cbd", "Cbg", "Cbs") if k in dev}
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), **extra,
            Cdd=dev["Cdd"], Cgd=dev["Cgd"], Cds=dev["Cds"], Cgs=dev["Cgs"], params=mb.params.view(np.uint8),
            scans=scans.view(np.uint8), algorithm=alg, flags=flags, z=z, n=n, margin=margin)
        print(name, z.shape, "n max", n.max(), "min margin", margin.min())


if __name__ == "__main__":
    main()
