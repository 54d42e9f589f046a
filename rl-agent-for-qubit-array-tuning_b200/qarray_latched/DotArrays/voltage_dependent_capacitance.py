"""Voltage-dependent capacitance holders (src/qarray_latched/DotArrays/voltage_dependent_capacitance.py:78-168).
The shipped configuration sets ``voltage_capacitance_model.type: null`` (qarray_config.yaml:134); with ``type: linear`` the
facade builds ``create_linear_capacitance_model(cdd_0=model.cdd_full, cgd_0=model.cgd_full, alpha, beta)``
(qarray_base_class.py:846-851).  Every factory of the reference is a pair of SCALAR scale factors per pixel, applied in the
ground state (ground_state.py:53-58) -- ``cgd = cgd_0 (1 + beta mean|v|)`` in all of them, and for ``cdd``:

    linear      cdd_0 (1 + alpha mean|v|)                         (:78-83, 123-135)
    quadratic   cdd_0 (1 + gamma sum v^2)                          (:94-99, 138-151)
    sigmoid     cdd_0 (1 + delta sigmoid(|v|_2 / v_char - 1))      (:102-109, 154-168)

so a model is a kind and three numbers; the tunnel kernels apply them (``qd_env_params.vc_kind / vc_alpha / vc_beta /
vc_vchar``).  ``alpha`` holds the cdd coefficient of the kind (alpha, gamma or delta)."""
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
    v_char: float = 1.0


def create_linear_capacitance_model(cdd_0, cgd_0, alpha: float = 0.1, beta: float = 0.01):
    return VoltageDependendentCapacitanceModel("linear", np.asarray(cdd_0), np.asarray(cgd_0), float(alpha), float(beta))


def create_quadratic_capacitance_model(cdd_0, cgd_0, gamma: float = 0.01, beta: float = 0.01):
    return VoltageDependendentCapacitanceModel("quadratic", np.asarray(cdd_0), np.asarray(cgd_0), float(gamma), float(beta))


def create_sigmoid_capacitance_model(cdd_0, cgd_0, v_char: float = 1.0, delta: float = 0.5, beta: float = 0.01):
    return VoltageDependendentCapacitanceModel("sigmoid", np.asarray(cdd_0), np.asarray(cgd_0), float(delta), float(beta),
                                               float(v_char))
