"""Host-side capacitance algebra, batched over envs (leading axes).  Runs once per env reset, never per pixel.

Mirrors the reference's Maxwell conversion (src/qarray_latched/DotArrays/_helper_functions.py:29-164):
``cdd = diag(rowsum(Cdd) + rowsum(Cgd)) - offdiag(Cdd)``, ``cgd = -Cgd``; the sensor enters as an extra row / column of
Cdd and an extra row of Cgd; barriers enter as extra *columns* of Cgd (voltage sources only).
"""
from __future__ import annotations

import numpy as np


def maxwell(cdd_nm, cgd_nm):
    """(..., n, n), (..., n, g) -> (cdd, cdd_inv, cgd) in Maxwell form."""
    cdd_nm = np.asarray(cdd_nm, dtype=np.float64)
    cgd_nm = np.asarray(cgd_nm, dtype=np.float64)
    n = cdd_nm.shape[-1]
    eye = np.eye(n, dtype=bool)
    total = cdd_nm.sum(axis=-1) + cgd_nm.sum(axis=-1)          # row sums taken before the diagonal is cleared
    off = np.where(eye, 0.0, cdd_nm)
    cdd = np.where(eye, total[..., :, None], 0.0) - off
    return cdd, np.linalg.inv(cdd), -cgd_nm


def embed_sensor(Cdd, Cgd, Cds, Cgs, Cbd=None, Cbs=None):
    """Non-Maxwell block matrices of the [dots, sensors] x [gates (, barriers)] system."""
    Cdd, Cgd, Cds, Cgs = (np.asarray(a, dtype=np.float64) for a in (Cdd, Cgd, Cds, Cgs))
    lead = Cdd.shape[:-2]
    n_dot, n_sensor, n_gate = Cdd.shape[-1], Cds.shape[-2], Cgd.shape[-1]
    n_barrier = 0 if Cbd is None else np.asarray(Cbd).shape[-1]
    d = n_dot + n_sensor
    cdd_full = np.zeros(lead + (d, d))
    cdd_full[..., :n_dot, :n_dot] = Cdd
    cdd_full[..., n_dot:, :n_dot] = Cds
    cdd_full[..., :n_dot, n_dot:] = np.swapaxes(Cds, -1, -2)
    cgd_full = np.zeros(lead + (d, n_gate + n_barrier))
    cgd_full[..., :n_dot, :n_gate] = Cgd
    cgd_full[..., n_dot:, :n_gate] = Cgs
    if n_barrier:
        cgd_full[..., :n_dot, n_gate:] = np.asarray(Cbd, dtype=np.float64)
        if Cbs is not None:
            cgd_full[..., n_dot:, n_gate:] = np.asarray(Cbs, dtype=np.float64)
    return cdd_full, cgd_full


def optimal_vg(cdd_inv, cgd, n_charges, rcond: float = 1e-3):
    """Voltages minimising the free energy of ``n_charges`` (TunnelCoupledChargeSensed.py:445-471), batched."""
    r = np.swapaxes(np.linalg.cholesky(np.asarray(cdd_inv, dtype=np.float64)), -1, -2)
    m = np.linalg.pinv(r @ np.asarray(cgd, dtype=np.float64), rcond=rcond) @ r
    return np.einsum("...ij,...j->...i", m, np.asarray(n_charges, dtype=np.float64))


def optimal_vgm(cdd_inv_full, cgd_gates, electrons: bool = False):
    """-pinv(cdd_inv_full @ cgd_full[:, :n_gate]) (TunnelCoupledChargeSensed.py:176-183)."""
    vgm = -np.linalg.pinv(np.asarray(cdd_inv_full) @ np.asarray(cgd_gates))
    return -vgm if electrons else vgm
