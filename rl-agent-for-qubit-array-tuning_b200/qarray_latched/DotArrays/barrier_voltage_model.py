"""``BarrierVoltageModel(n_barrier, n_dot, tc_base, alpha)`` with the reference's constructor and attributes
(src/qarray_latched/DotArrays/barrier_voltage_model.py:21-53, 194-216; constructed at
src/qadapt/environment/qarray_base_class.py:806-811, ``.alpha`` / ``.tc_base`` read at :813-814).

A parameter holder plus small host-side evaluators; per-pixel tunnel couplings are computed inside the CUDA kernel as
``tc_d = tc_base * exp(-alpha_d * (vb_d + (Cbg vg)_d))`` (reference :83, :129; the cross-barrier term of :142 is the
diagonal of a zero-diagonal matrix and contributes nothing)."""
from __future__ import annotations

import numpy as np


class BarrierVoltageModel:
    def __init__(self, n_barrier: int, n_dot: int, tc_base: float = 0.1, alpha=1.0):
        self.n_barrier = int(n_barrier)
        self.n_dot = int(n_dot)
        self.tc_base = tc_base
        if isinstance(alpha, (list, tuple, np.ndarray)):
            if len(alpha) != n_barrier:
                raise ValueError(f"Alpha list length ({len(alpha)}) must match n_barrier ({n_barrier})")
            self.alpha = np.asarray(alpha, dtype=np.float64)
        else:
            self.alpha = np.full(n_barrier, float(alpha))

    def compute_effective_barrier_potential(self, vg, vb, parent_model):
        v = np.asarray(vb, dtype=np.float64)
        if parent_model.Cbg is not None:
            v = v + np.einsum("bg,...g->...b", np.asarray(parent_model.Cbg), np.asarray(vg, dtype=np.float64))
        return v

    def compute_tc_matrix(self, vb_eff):
        nb = self.n_dot - 1
        if self.n_barrier < nb:
            raise ValueError(f"Linear topology with {self.n_dot} dots requires {nb} barriers, got {self.n_barrier}")
        t = self.tc_base * np.exp(-self.alpha[:nb] * np.asarray(vb_eff, dtype=np.float64)[..., :nb])
        out = np.zeros(t.shape[:-1] + (self.n_dot, self.n_dot))
        i = np.arange(nb)
        out[..., i, i + 1] = t
        out[..., i + 1, i] = t
        return out

    compute_tc_matrix_batch = compute_tc_matrix

    def validate_dimensions(self, n_gate: int, n_dot: int, n_sensor: int):
        if self.n_dot != n_dot:
            raise ValueError(f"Barrier model n_dot={self.n_dot} doesn't match parent n_dot={n_dot}")
        if self.n_barrier < 1:
            raise ValueError("Must have at least 1 barrier")
        if self.n_barrier < self.n_dot - 1:
            raise ValueError(f"Linear topology with {self.n_dot} dots requires {self.n_dot - 1} barriers, got {self.n_barrier}")
        if len(self.alpha) != self.n_barrier:
            raise ValueError(f"Alpha array length ({len(self.alpha)}) must match n_barrier ({self.n_barrier})")
