"""``LatchingModel(n_dots, p_leads, p_inter)`` as constructed by the reference
(src/qadapt/environment/qarray_base_class.py:732-737, 495-519).  A parameter holder: the sequential latching pass runs in
the CUDA kernel (one warp per scan row, see csrc/qd_kernels.cuh)."""
from __future__ import annotations

import numpy as np


class LatchingBaseModel:
    """No latching."""
    exists = False

    def add_latching(self, n, measurement_shape=None):
        return n


class LatchingModel(LatchingBaseModel):
    exists = True

    def __init__(self, n_dots: int, p_leads, p_inter):
        self.n_dots = int(n_dots)
        p_leads = np.asarray(p_leads, dtype=np.float64)
        p_inter = np.asarray(p_inter, dtype=np.float64)
        if p_leads.ndim == 0:
            p_leads = np.full(self.n_dots, float(p_leads))
        if p_inter.ndim == 0:
            p_inter = np.full((self.n_dots, self.n_dots), float(p_inter))
            np.fill_diagonal(p_inter, 0.0)
        assert p_leads.shape == (self.n_dots,), "p_leads must be of shape (n_dots,)"
        assert p_inter.shape == (self.n_dots, self.n_dots), "p_inter must be of shape (n_dots, n_dots)"
        assert np.allclose(p_inter, p_inter.T), "p_inter must be symmetric"
        self.p_leads = p_leads
        self.p_inter = p_inter

    def add_latching(self, n, measurement_shape=None):
        raise NotImplementedError("latching is fused into the CUDA scan kernel; call do2d_open / charge_sensor_open")
