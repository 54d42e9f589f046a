"""Hysteresis latching along the fast scan axis.

Test infrastructure (see ``oracle/__init__.py``).  PARITY UNPINNED: restates qarray==1.6.0 ``LatchingModel``
(absent).  Anchors: construction ``LatchingModel(n_dots, p_leads (N,), p_inter (N,N))`` at
src/qadapt/environment/qarray_base_class.py:732-737 and :495-519; invocation ``add_latching(n, measurement_shape)`` at
src/qarray_latched/DotArrays/ground_state.py:164; the same author's in-tree prototype
src/qarray_latched/latched.py:86-159 (hold the whole previous state when a transition is rejected, :114-120; accept
unconditionally at the first pixel of each row, :155).

Rule (per pixel, sequential along x): let ``h`` be the held configuration and ``c`` the unlatched ground state here.
Count the dots where they differ:  0 -> ``h`` (= ``c``);  1 (dot i) -> accept ``c`` with probability ``p_leads[i]``;
2 (dots i<j) -> accept with probability ``p_inter[i, j]``;  >2 -> accept;  rejected -> keep ``h`` whole.
"accept with probability p" is ``u < p`` with ``u`` the pixel's latching uniform (``oracle.philox``).

Switches:
* ``compare``: ``"rounded"`` (default) compares ``floor(n + 1/2)``; ``"exact"`` compares the floats elementwise.
* ``carry_rows``: False (default) resets ``h`` to ``c`` at ix == 0 of every row; True runs one flat pass over the scan.
"""
from __future__ import annotations

import numpy as np


def add_latching(n, u_latch, p_leads, p_inter, compare: str = "rounded", carry_rows: bool = False):
    """``n`` (ny, nx, N) float, ``u_latch`` (ny, nx) -> latched copy of ``n``."""
    n = np.asarray(n, dtype=np.float64)
    ny, nx, n_dot = n.shape
    p_leads = np.asarray(p_leads, dtype=np.float64)
    p_inter = np.asarray(p_inter, dtype=np.float64)
    key = np.floor(n + 0.5) if compare == "rounded" else n
    out = n.copy()
    held = None
    held_key = None
    for iy in range(ny):
        for ix in range(nx):
            c, ck = n[iy, ix], key[iy, ix]
            if held is None or (ix == 0 and not carry_rows):
                held, held_key = c, ck
                continue
            diff = np.nonzero(ck != held_key)[0]
            if diff.size == 0:
                accept = True
            elif diff.size == 1:
                accept = u_latch[iy, ix] < p_leads[diff[0]]
            elif diff.size == 2:
                accept = u_latch[iy, ix] < p_inter[diff[0], diff[1]]
            else:
                accept = True
            if accept:
                held, held_key = c, ck
            out[iy, ix] = held
    return out
