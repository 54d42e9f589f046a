"""One whole scan, CPU side: ground state -> latching -> sensor -> noise, for one env and one scan window.

Test infrastructure (see ``oracle/__init__.py``).  This is the composition the reference performs per adjacent dot pair
in ``QarrayBaseClass._get_charge_sensor_data`` (src/qadapt/environment/qarray_base_class.py:95-168) followed by
``_apply_radial_noise`` (:202-206, 444-493), written against the same inputs the CUDA kernel takes (an affine scan
descriptor + per-env matrices) so the two can be compared pixel for pixel.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import composer, latching, noise, path_a, philox, sensor


@dataclass
class Model:
    """Per-env constants (Maxwell form, fp64)."""
    cdd_inv: np.ndarray            # (N, N)  matrix of the ground-state quadratic form
    cdd: np.ndarray                # (N, N)  its inverse (M-matrix) -- used by the exact relaxation (Path A)
    cgd: np.ndarray                # (N, n_volt)
    cdd_inv_full: np.ndarray       # (D, D)
    cgd_full: np.ndarray           # (D, n_volt)
    algorithm: str = "default"     # default | thresholded | brute_force | tunnel (Path B)
    threshold: float = 1.0
    max_charge_carriers: int = 4
    kT: float = 0.0
    latching: bool = False
    p_leads: np.ndarray | None = None
    p_inter: np.ndarray | None = None
    white_amp: float = 0.0
    tele_p01: float = 0.0
    tele_p10: float = 0.0
    tele_amp: float = 0.0
    pink_amp: float = 0.0          # 1/f input noise (oracle/noise.py::pink_noise); 0 = off
    # Path B only
    n_gate: int = 0
    cbg: np.ndarray | None = None  # (B, G) raw positive
    tc_base: float = 0.0
    alpha: np.ndarray | None = None
    num_charge_states: int = 32
    charge_state_batch_size: int = 1000
    vc_alpha: float = 0.0          # create_linear_capacitance_model(alpha, beta); 0, 0 = constant capacitances
    vc_beta: float = 0.0
    vc_kind: int = 0               # 0 linear, 1 quadratic (vc_alpha = gamma), 2 sigmoid (vc_alpha = delta, vc_vchar = v_char)
    vc_vchar: float = 1.0


@dataclass
class Scan:
    v0: np.ndarray
    dx: np.ndarray
    dy: np.ndarray
    nx: int
    ny: int
    peak_width: float
    seed: int = 0
    rad_mode: int = 0
    rad: tuple = field(default_factory=lambda: (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0))  # x0, dx, y0, dy, alpha, zero_r, max_amp


def simulate_scan(m: Model, s: Scan, latch_compare: str = "rounded", carry_rows: bool = False,
                  white_on: str = "input", return_margin: bool = False):
    """Returns ``(z (ny, nx) float64, n (ny, nx, N) float64[, margin (ny, nx)])``."""
    v = composer.affine_grid(s.v0, s.dx, s.dy, s.nx, s.ny)
    return simulate_points(m, v, s, latch_compare, carry_rows, white_on, return_margin)


def simulate_points(m: Model, v, s: Scan, latch_compare: str = "rounded", carry_rows: bool = False,
                    white_on: str = "input", return_margin: bool = False):
    """Same on an explicit voltage array ``v`` (ny, nx, n_volt); ``s`` supplies peak width, seed and radial fields."""
    v = np.asarray(v, dtype=np.float64)
    ny, nx = v.shape[:2]
    n_dot = m.cdd_inv.shape[0]
    v = v.reshape(ny * nx, -1)
    if m.algorithm == "tunnel":
        from . import path_b
        n, margin = path_b.ground_state_open(m, v, return_gap=True)     # "margin" = spectral gap of the pixel
    else:
        n, margin = path_a.ground_state_open(v, m.cgd, m.cdd_inv, m.cdd, m.algorithm, m.threshold,
                                             m.max_charge_carriers, m.kT, return_margin=True)
    n = n.reshape(ny, nx, n_dot)
    draws = philox.pixel_draws(s.seed, ny * nx)
    u_row = philox.row_draws(s.seed, ny)
    if m.latching:
        n = latching.add_latching(n, draws["u_latch"].reshape(ny, nx), m.p_leads, m.p_inter,
                                  compare=latch_compare, carry_rows=carry_rows)
    noise_in = noise.input_noise(draws, ny, nx, m.white_amp, m.tele_p01, m.tele_p10, m.tele_amp, u_row,
                                 carry_rows=carry_rows, white_on=white_on)
    if m.pink_amp != 0.0:
        noise_in = noise_in + m.pink_amp * noise.pink_noise(s.seed, ny, nx, carry_rows=carry_rows)
    out_noise = None
    if white_on == "output" and m.white_amp != 0.0:
        out_noise = (m.white_amp * draws["z_white"]).reshape(ny * nx, 1)
    z = sensor.charge_sensor_signal(n.reshape(ny * nx, n_dot), v, m.cdd_inv_full, m.cgd_full, s.peak_width,
                                    input_noise=noise_in.reshape(ny * nx, 1), output_noise=out_noise)
    z = z.reshape(ny, nx)
    z = noise.radial_noise(z, draws["z_radial"], s.rad_mode, *s.rad)
    if return_margin:
        return z, n, margin.reshape(ny, nx)
    return z, n
