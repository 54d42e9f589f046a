"""Batched observation stage: what ``QarrayBaseClass._get_obs`` + ``QuantumDeviceEnv._normalise_obs`` do for ONE env
(src/qadapt/environment/qarray_base_class.py:95-229, 444-493; env.py:471-509), for a whole batch of envs at once.

``obs_scans`` turns the batch's env state (plunger / barrier / sensor voltages, per-env virtual gate matrix and origin,
ground truth for the radial noise) into the N-1 scan descriptors per env -- vectorised NumPy, no per-env Python -- and
``observe`` launches them and normalises the images on the device.  Output layout ``[env, pair, iy, ix]`` float32.
"""
from __future__ import annotations

import numpy as np

from ._lib import FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL, N_NONE
from .engine import Engine, ModelBatch, new_scans


def obs_scans(mb: ModelBatch, gate_voltages, sensor_voltage, vgm, origin, obs_min, obs_max, res: int,
              barrier_voltages=None, peak_width=0.1, peak_width_alpha=None, virtual: bool = True,
              gate_ground_truth=None, radial: dict | None = None, seeds=None):
    """Scan descriptors of one env.step for every env (env-major, N-1 adjacent pairs each).

    gate_voltages (E, N); sensor_voltage scalar or (E,); vgm (E, G, G); origin (E, G); barrier_voltages (E, B) for the
    tunnel path; obs_min / obs_max scalar or (E,) (the env draws its window half-width per episode, env.py:160-172).  ``virtual=True`` is the facade's barrier-mode call (``do2d('vP{i}', ..., 'vP{i+1}', ...,
    gate_voltages, add_full_crosstalk=True)``, :143-154); ``virtual=False`` its non-barrier call (physical gates i, i+1,
    every other gate at 0, :128-137).  ``radial``: dict(zero_radius (E,), ramp_distance (E,), full_noise_distance (E,) or
    None, max_amplitude) as sampled at :409-441.  ``peak_width_alpha``: VaryPeakWidth (utils/vary_peak_width.py).
    """
    gv = np.asarray(gate_voltages, dtype=np.float64)
    E, N = gv.shape
    G = mb.n_gate
    assert E == mb.n_env and N == mb.n_dot and G == N + 1
    n_scan = E * (N - 1)
    env = np.repeat(np.arange(E), N - 1)
    ch = np.tile(np.arange(N - 1), E)
    rows = np.arange(n_scan)
    obs_min = np.broadcast_to(np.asarray(obs_min, dtype=np.float64), (E,))[env]       # scalar or per env (window_delta)
    obs_max = np.broadcast_to(np.asarray(obs_max, dtype=np.float64), (E,))[env]
    step = (obs_max - obs_min) / (res - 1) if res > 1 else np.zeros(n_scan)
    scans = new_scans(n_scan)
    v1, v2 = gv[env, ch], gv[env, ch + 1]
    if virtual:
        vgm = np.broadcast_to(np.asarray(vgm, dtype=np.float64), (E, G, G))
        origin = np.broadcast_to(np.asarray(origin, dtype=np.float64), (E, G))
        sv = np.broadcast_to(np.asarray(sensor_voltage, dtype=np.float64), (E,))
        base = np.concatenate([gv, sv[:, None]], axis=1)[env]                  # (n_scan, G): all dots at their voltage
        base[rows, ch] = v1 + obs_min
        base[rows, ch + 1] = v2 + obs_min
        scans["v0"][:, :G] = np.einsum("sij,sj->si", vgm[env], base) + origin[env]
        scans["dx"][:, :G] = vgm[env, :, ch] * step[:, None]
        scans["dy"][:, :G] = vgm[env, :, ch + 1] * step[:, None]
    else:
        scans["v0"][rows, ch] = v1 + obs_min
        scans["v0"][rows, ch + 1] = v2 + obs_min
        scans["dx"][rows, ch] = step
        scans["dy"][rows, ch + 1] = step
    if mb.n_volt > G:
        assert barrier_voltages is not None, "Barrier voltages must be provided for models with barriers"
        scans["v0"][:, G:mb.n_volt] = np.asarray(barrier_voltages, dtype=np.float64)[env]
    pw = np.broadcast_to(np.asarray(peak_width, dtype=np.float64), (E,))[env]
    if peak_width_alpha is not None:
        alpha = np.broadcast_to(np.asarray(peak_width_alpha, dtype=np.float64), (E,))[env]
        pw = np.clip(pw - np.abs(alpha * (np.abs(v1) + np.abs(v2)) / 2), 0, 1)
    scans["peak_width"] = pw
    scans["env_id"], scans["nx"], scans["ny"] = env, res, res
    scans["pix_offset"] = rows.astype(np.int64) * res * res
    scans["seed"] = (np.random.randint(0, 2 ** 62, size=n_scan, dtype=np.int64).astype(np.uint64)
                     if seeds is None else np.asarray(seeds, dtype=np.uint64))
    if radial is not None and gate_ground_truth is not None:
        gt = np.asarray(gate_ground_truth, dtype=np.float64)
        gt1, gt2 = gt[env, ch], gt[env, ch + 1]
        zero_r = np.broadcast_to(np.asarray(radial["zero_radius"], dtype=np.float64), (E,))[env]
        ramp = np.broadcast_to(np.asarray(radial["ramp_distance"], dtype=np.float64), (E,))[env]
        scans["rad_mode"] = 1
        scans["rad_x0"], scans["rad_dx"] = v1 + obs_min - gt1, step
        scans["rad_y0"], scans["rad_dy"] = v2 + obs_min - gt2, step
        scans["rad_max_amp"] = radial["max_amplitude"]
        scans["rad_alpha"] = radial["max_amplitude"] / ramp
        scans["rad_zero_radius"] = zero_r
        full = radial.get("full_noise_distance")
        if full is not None:
            full = np.broadcast_to(np.asarray(full, dtype=np.float64), (E,))[env]
            scans["rad_mode"][(np.abs(v1 - gt1) > full) | (np.abs(v2 - gt2) > full)] = 2
    return scans


def observe(eng: Engine, scans, z_dev, flags: int = FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL, normalise: bool = True,
            stream=None):
    """Launch the batch and (optionally) normalise per env on the device.  ``z_dev``: CUDA float32 tensor with
    ``n_env * (N-1) * res * res`` elements; returns it viewed as ``[env, pair, iy, ix]``."""
    mb = eng.models
    eng.scan_open(scans, z_dev, None, N_NONE, flags, stream)
    if normalise:
        eng.normalise_obs(z_dev, stream=stream)
    res_y, res_x = int(scans["ny"][0]), int(scans["nx"][0])
    return z_dev.view(mb.n_env, mb.n_dot - 1, res_y, res_x)
