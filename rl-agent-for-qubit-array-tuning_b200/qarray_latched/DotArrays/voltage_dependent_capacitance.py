"""Voltage-dependent capacitance holder (src/qarray_latched/DotArrays/voltage_dependent_capacitance.py:78-168).
The shipped configuration sets ``voltage_capacitance_model.type: null`` (qarray_config.yaml:134); with ``type: linear`` the
facade builds ``create_linear_capacitance_model(cdd_0=model.cdd_full, cgd_0=model.cgd_full, alpha, beta)``
(qarray_base_class.py:846-851).  The model is two scalars: per pixel ``cdd = cdd_0 (1 + alpha mean|v|)``,
``cgd = cgd_0 (1 + beta mean|v|)`` in the ground state (ground_state.py:53-58); the tunnel kernel applies them
(``qd_env_params.vc_alpha / vc_beta``)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class VoltageDependendentCapacitanceModel:
    kind: str
    cdd_0: np.ndarray
    cgd_0: np.ndarray
    alpha: float
    beta: float


def create_linear_capacitance_model(cdd_0, cgd_0, alpha: float = 0.1, beta: float = 0.01):
    return VoltageDependendentCapacitanceModel("linear", np.asarray(cdd_0), np.asarray(cgd_0), float(alpha), float(beta))
