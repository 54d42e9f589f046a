"""Voltage-dependent capacitance holder (src/qarray_latched/DotArrays/voltage_dependent_capacitance.py:78-168).
The shipped configuration sets ``voltage_capacitance_model.type: null`` (qarray_config.yaml:134), so the reference never
builds one on the hot path; the linear model is accepted and stored, and the kernel rejects it loudly until its per-pixel
scaling ``1 + alpha * mean(abs(v))`` is wired in."""
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
