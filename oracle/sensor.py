"""Charge-sensor response.

Test infrastructure (see ``oracle/__init__.py``).  Literal restatement of
src/qarray_latched/DotArrays/TunnelCoupledChargeSensed.py:320-380 (``charge_sensor_open``; the same stack without
``vb`` is the upstream ``ChargeSensedDotArray.charge_sensor_open``, cf. the untouched upstream-style copy at :391-426)
and of ``lorentzian`` src/qarray_latched/DotArrays/_helper_functions.py:167-177.  ``constant_charge_shift`` is None in
every reference call site and is not restated.
"""
from __future__ import annotations

import numpy as np


def lorentzian(x, x0, gamma):
    return np.reciprocal(((x - x0) / gamma) ** 2 + 1)


def charge_sensor_signal(n_open, v_ext, cdd_inv_full, cgd_full, coulomb_peak_width, n_peak: int = 5,
                         input_noise=None, output_noise=None):
    """``n_open`` (..., N), ``v_ext`` (..., n_volt) -> signal (..., n_sensor).

    Eleven full-system free energies with the sensor occupation perturbed by k = -n_peak..n_peak, first differences,
    ten Lorentzians, summed.
    """
    n_open = np.asarray(n_open, dtype=np.float64)
    v_ext = np.asarray(v_ext, dtype=np.float64)
    cdd_inv_full = np.asarray(cdd_inv_full, dtype=np.float64)
    cgd_full = np.asarray(cgd_full, dtype=np.float64)
    n_dot = n_open.shape[-1]
    n_sensor = cdd_inv_full.shape[0] - n_dot
    n_cont = np.einsum("ij,...j", cgd_full, v_ext)
    n_sens = np.round(n_cont[..., n_dot:n_dot + n_sensor])
    if input_noise is None:
        input_noise = np.zeros(n_sens.shape)
    if output_noise is None:
        output_noise = np.zeros(n_sens.shape)
    f = np.zeros((2 * n_peak + 1, *n_sens.shape))
    v_dash = np.einsum("ij,...j", cgd_full, v_ext)
    for sensor in range(n_sensor):
        for i, k in enumerate(range(-n_peak, n_peak + 1)):
            pert = n_sens.copy()
            pert[..., sensor] = pert[..., sensor] + k
            n_full = np.concatenate([n_open, pert + input_noise], axis=-1)
            d = n_full - v_dash
            f[i, ..., sensor] = np.einsum("...i,ij,...j", d, cdd_inv_full, d)
    signal = lorentzian(np.diff(f, axis=0), 0, coulomb_peak_width).sum(axis=0)
    return signal + output_noise
