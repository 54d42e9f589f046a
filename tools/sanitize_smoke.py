#!/usr/bin/env python
"""Tiny end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Path A default / thresholded / brute force / thermal, latching + noise, row and flat passes, points mode, Path B."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rl-agent-for-qubit-array-tuning_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

from qdsim import (FLAG_CARRY_ROWS, FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL, FLAG_THERMAL, N_F64, N_U8, Engine,  # noqa: E402
                   synth)

eng = Engine(0)
full = FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL
for n_dot, alg, flags, kw in ((8, "default", full, {}), (5, "thresholded", full | FLAG_CARRY_ROWS, {}),
                              (3, "brute_force", FLAG_LATCH, {}), (4, "default", FLAG_THERMAL | FLAG_LATCH, {"thermal": True})):
    dev = synth.sample_devices(3, n_dot, seed=1)
    mb = synth.model_batch(dev, algorithm=alg, **kw)
    eng.set_models(mb)
    sc = synth.env_step_scans(mb, dev, res=19, seed=2, offset_range=3.0)
    z, n = eng.scan_open_host(sc, n_type=N_F64 if flags & FLAG_THERMAL else N_U8, flags=flags)
    assert np.isfinite(z).all()
    v = np.random.default_rng(0).uniform(-3, 1, (5, 37, mb.n_volt))
    eng.points_open_host(sc[0], v, n_type=N_F64, flags=flags & ~FLAG_RADIAL)
for n_dot in (4, 6):
    dev = synth.sample_barrier_devices(2, n_dot, seed=3)
    mb = synth.tunnel_batch(dev)
    eng.set_models(mb)
    sc = synth.env_step_scans(mb, dev, res=9, seed=4, offset_range=2.0)
    z, n = eng.scan_open_host(sc, n_type=N_F64, flags=full)
    assert np.isfinite(z).all() and np.isfinite(n).all()
eng.close()
print("sanitize_smoke ok")
