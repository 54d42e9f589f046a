"""Sensor noise models.

Test infrastructure (see ``oracle/__init__.py``).

* White + telegraph noise: PARITY UNPINNED restatement of qarray==1.6.0 ``WhiteNoise(amplitude)``,
  ``TelegraphNoise(p01, p10, amplitude)`` and their sum (absent).  Anchors: construction
  src/qadapt/environment/qarray_base_class.py:726-728; consumption as ``sample_input_noise`` (added to the sensor
  occupation before the Lorentzian) and ``sample_output_noise`` (added to the signal) at
  src/qarray_latched/DotArrays/TunnelCoupledChargeSensed.py:354, 379; parameter ranges qarray_config.yaml:48-54 and
  ``p10 = factor * p01`` qarray_base_class.py:398-401, 436.
    - white: iid ``N(0, amplitude^2)``; injected as input noise (switch ``white_on``: "input" | "output").
    - telegraph: two-state Markov chain along the flattened pixels; from state 0 it flips with probability ``p01`` per
      pixel, from state 1 with ``p10``; contributes ``amplitude`` while in state 1; injected as input noise.
      ``carry_rows=True``: one chain per scan starting in state 0.  ``carry_rows=False`` (default, rows independent):
      each row's chain starts from the stationary distribution, state 1 with probability ``p01/(p01+p10)``.
* Radial noise: literal restatement of src/qadapt/environment/qarray_base_class.py:444-493 (in tree).
"""
from __future__ import annotations

import numpy as np


def telegraph_states(u_tele, u_row, p01: float, p10: float, carry_rows: bool = False):
    """``u_tele`` (ny, nx) per-pixel uniforms, ``u_row`` (ny,) -> int states (ny, nx)."""
    ny, nx = u_tele.shape
    out = np.zeros((ny, nx), dtype=np.int64)
    s = 0
    tot = p01 + p10
    for iy in range(ny):
        if not carry_rows:
            s = 1 if (tot > 0.0 and u_row[iy] < p01 / tot) else 0
        for ix in range(nx):
            p = p10 if s else p01
            if u_tele[iy, ix] < p:
                s ^= 1
            out[iy, ix] = s
    return out


def input_noise(draws, ny, nx, white_amp, p01, p10, tele_amp, u_row, carry_rows=False, white_on="input"):
    """Sum of the white and telegraph input-noise fields, shape (ny, nx)."""
    z = np.zeros((ny, nx))
    if white_on == "input" and white_amp != 0.0:
        z = z + white_amp * draws["z_white"].reshape(ny, nx)
    if tele_amp != 0.0:
        z = z + tele_amp * telegraph_states(draws["u_tele"].reshape(ny, nx), u_row, p01, p10, carry_rows)
    return z


PINK_TAUS = (2.0, 8.0, 32.0, 128.0)


def pink_noise(seed: int, ny: int, nx: int, carry_rows: bool = False):
    """1/f sensor input noise, unit variance, shape (ny, nx) -- ``north_star`` names it; the reference has NO 1/f model
    (SURVEY.md 8a S6), so this is the framework's own definition, restated here for the CUDA kernel to be checked against:

    the sum of four Ornstein-Uhlenbeck chains along the fast axis, ``x_k[i] = a_k x_k[i-1] + sqrt(1 - a_k^2) xi_k[i]`` with
    ``a_k = exp(-1 / tau_k)``, ``tau_k = 2, 8, 32, 128`` pixels (each of unit stationary variance; their sum / 2 has a spectrum
    close to 1/f between 1/128 and 1/2 cycles per pixel).  ``xi_k[i]``: the four Box-Muller normals of Philox stream 2 at the
    pixel; the state before the first pixel of a row (of the scan when ``carry_rows``): the same four normals of stream 3 at
    the row index."""
    from . import philox

    def four_normals(index, purpose):
        w0, w1, w2, w3 = philox.words(seed, index, purpose)
        out = []
        for wa, wb in ((w0, w1), (w2, w3)):
            u1 = ((wa >> np.uint32(8)).astype(np.float64) + 1.0) * philox.INV_2_24
            u2 = philox.u24(wb)
            r = np.sqrt(-2.0 * np.log(u1))
            out += [r * np.cos(2.0 * np.pi * u2), r * np.sin(2.0 * np.pi * u2)]
        return np.stack(out, axis=-1)                      # (..., 4)

    xi = four_normals(np.arange(ny * nx, dtype=np.uint64), 2).reshape(ny, nx, 4)
    start = four_normals(np.arange(ny, dtype=np.uint64), 3)            # (ny, 4)
    a = np.exp(-1.0 / np.asarray(PINK_TAUS))
    b = np.sqrt(1.0 - a * a)
    out = np.zeros((ny, nx))
    x = start[0].copy()
    for iy in range(ny):
        if not carry_rows:
            x = start[iy].copy()
        for ix in range(nx):
            x = a * x + b * xi[iy, ix]
            out[iy, ix] = 0.5 * x.sum()
    return out


def radial_noise(z, z_radial, mode: int, x0, dx, y0, dy, alpha, zero_radius, max_amplitude):
    """``mode`` 0: off; 1: ``z + randn * clip(alpha*(dist - zero_radius), 0, max_amplitude)``; 2: ``randn`` replaces z.

    ``dist[iy, ix] = hypot(x0 + ix*dx, y0 + iy*dy)`` -- distance of the pixel from the ground-truth voltages
    (qarray_base_class.py:476-486: ``x0 = v1 + obs_voltage_min - gt1``, ``dx = (max-min)/(res-1)``).
    """
    ny, nx = z.shape
    zr = z_radial.reshape(ny, nx)
    if mode == 0:
        return z
    if mode == 2:
        return zr.copy()
    vx = x0 + np.arange(nx) * dx
    vy = y0 + np.arange(ny) * dy
    dist = np.sqrt(vx[None, :] ** 2 + vy[:, None] ** 2)
    amp = np.clip(alpha * (dist - zero_radius), 0.0, max_amplitude)
    return z + zr * amp
