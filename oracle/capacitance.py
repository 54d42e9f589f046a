"""Maxwell conversion, optimal gate voltages and virtual-gate matrices.

Test infrastructure (see ``oracle/__init__.py``).  Restates, in plain NumPy fp64:

* ``convert_to_maxwell``                 <- src/qarray_latched/DotArrays/_helper_functions.py:129-164
* ``with_sensor``                        <- src/qarray_latched/DotArrays/_helper_functions.py:29-57
* ``with_barriers_and_sensor``           <- src/qarray_latched/DotArrays/_helper_functions.py:60-126
* ``optimal_vg``                         <- src/qarray_latched/DotArrays/TunnelCoupledChargeSensed.py:445-471
                                            and src/qarray_latched/optimal_v_calc.py:10-22
* ``optimal_virtual_gate_matrix``        <- src/qarray_latched/DotArrays/TunnelCoupledChargeSensed.py:176-183
                                            (the line the reference itself cites for upstream at
                                            src/qadapt/environment/qarray_base_class.py:893-895)
"""
from __future__ import annotations

import numpy as np


def convert_to_maxwell(cdd_non_maxwell, cgd_non_maxwell):
    """cdd = diag(rowsum(Cdd) + rowsum(Cgd)) - offdiag(Cdd);  cdd_inv = inv(cdd);  cgd = -Cgd."""
    cdd_nm = np.array(cdd_non_maxwell, dtype=np.float64, copy=True)
    cgd_nm = np.array(cgd_non_maxwell, dtype=np.float64, copy=True)
    cdd_sum = cdd_nm.sum(axis=1)
    cgd_sum = cgd_nm.sum(axis=1)
    np.fill_diagonal(cdd_nm, 0.0)
    cdd = np.diag(cdd_sum + cgd_sum) - cdd_nm
    return cdd, np.linalg.inv(cdd), -cgd_nm


def with_sensor(Cdd, Cgd, Cds, Cgs):
    """Embed the sensor as extra row/column of Cdd and extra row of Cgd, then convert."""
    Cdd, Cgd, Cds, Cgs = (np.asarray(a, dtype=np.float64) for a in (Cdd, Cgd, Cds, Cgs))
    n_dot, n_sensor, n_gate = Cdd.shape[0], Cds.shape[0], Cgd.shape[1]
    cdd_full = np.zeros((n_dot + n_sensor, n_dot + n_sensor))
    cdd_full[:n_dot, :n_dot] = Cdd
    cdd_full[n_dot:, :n_dot] = Cds
    cdd_full[:n_dot, n_dot:] = Cds.T
    cgd_full = np.zeros((n_dot + n_sensor, n_gate))
    cgd_full[:n_dot] = Cgd
    cgd_full[n_dot:] = Cgs
    return convert_to_maxwell(cdd_full, cgd_full)


def with_barriers_and_sensor(Cdd, Cgd, Cds, Cgs, Cbd=None, Cbs=None):
    """Barriers are voltage sources: extra *columns* of Cgd (Cbd for dots, Cbs for sensors). Cbg/Cbb unused here."""
    Cdd, Cgd, Cds, Cgs = (np.asarray(a, dtype=np.float64) for a in (Cdd, Cgd, Cds, Cgs))
    n_dot, n_sensor, n_gate = Cdd.shape[0], Cds.shape[0], Cgd.shape[1]
    n_barrier = 0 if Cbd is None else np.asarray(Cbd).shape[1]
    d = n_dot + n_sensor
    cdd_full = np.zeros((d, d))
    cdd_full[:n_dot, :n_dot] = Cdd
    cdd_full[n_dot:, :n_dot] = Cds
    cdd_full[:n_dot, n_dot:] = Cds.T
    cgd_full = np.zeros((d, n_gate + n_barrier))
    cgd_full[:n_dot, :n_gate] = Cgd
    cgd_full[n_dot:, :n_gate] = Cgs
    if n_barrier:
        cgd_full[:n_dot, n_gate:] = np.asarray(Cbd, dtype=np.float64)
        if Cbs is not None:
            cgd_full[n_dot:, n_gate:] = np.asarray(Cbs, dtype=np.float64)
    return convert_to_maxwell(cdd_full, cgd_full)


def optimal_vg(cdd_inv, cgd, n_charges, rcond: float = 1e-3):
    """R = chol(cdd_inv)^T;  M = pinv(R cgd, rcond) R;  v = M n."""
    R = np.linalg.cholesky(np.asarray(cdd_inv, dtype=np.float64)).T
    M = np.linalg.pinv(R @ np.asarray(cgd, dtype=np.float64), rcond=rcond) @ R
    return np.einsum("ij,...j", M, np.asarray(n_charges, dtype=np.float64))


def optimal_virtual_gate_matrix(cdd_inv_full, cgd_gates_only, charge_carrier: str = "h"):
    """VGM = -pinv(cdd_inv_full @ cgd_full[:, :n_gate]); sign flipped for electrons."""
    vgm = -np.linalg.pinv(np.asarray(cdd_inv_full) @ np.asarray(cgd_gates_only))
    return -vgm if charge_carrier == "electrons" else vgm
