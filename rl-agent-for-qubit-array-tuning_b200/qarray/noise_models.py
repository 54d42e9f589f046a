"""Noise-model classes with the names and constructor signatures of qarray==1.6.0's ``qarray.noise_models`` as the
reference uses them (src/qadapt/environment/qarray_base_class.py:726-728: ``WhiteNoise(amplitude=...)``,
``TelegraphNoise(p01=..., p10=..., amplitude=...)``, ``white + telegraph``).

They are parameter holders: the draws happen inside the CUDA kernel from a counter-based Philox stream.  The
``sample_input_noise`` / ``sample_output_noise`` hooks (consumed at TunnelCoupledChargeSensed.py:354, 379 in the
reference) are kept for code that calls them directly; they report what the kernel will inject, as zeros, because the
kernel adds the noise itself.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


class BaseNoiseModel:
    """Adds no noise."""

    def sample_input_noise(self, shape):
        return np.zeros(shape)

    def sample_output_noise(self, shape):
        return np.zeros(shape)

    def __add__(self, other):
        return NoiseModelSum([self, other])

    # parameters seen by the kernel
    def _kernel_params(self):
        return {}


@dataclass
class WhiteNoise(BaseNoiseModel):
    amplitude: float = 0.0

    def _kernel_params(self):
        return {"white_amp": float(self.amplitude)}


@dataclass
class TelegraphNoise(BaseNoiseModel):
    p01: float = 0.0
    p10: float = 0.0
    amplitude: float = 0.0

    def _kernel_params(self):
        return {"tele_p01": float(self.p01), "tele_p10": float(self.p10), "tele_amp": float(self.amplitude)}


class NoiseModelSum(BaseNoiseModel):
    def __init__(self, models):
        self.models = []
        for m in models:
            self.models.extend(m.models if isinstance(m, NoiseModelSum) else [m])
        kinds = [type(m) for m in self.models if not type(m) is BaseNoiseModel]
        if len(kinds) != len(set(kinds)):
            raise NotImplementedError("the CUDA kernel carries one white and one telegraph source per device")

    def _kernel_params(self):
        out = {}
        for m in self.models:
            out.update(m._kernel_params())
        return out
