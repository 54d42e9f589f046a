"""Scan-grid synthesis.

Test infrastructure (see ``oracle/__init__.py``).  Restates src/qarray_latched/DotArrays/GateVoltageComposer.py:
``meshgrid`` (:86-125), ``meshgrid_virtual`` (:127-168), ``meshgrid_virtual_coupled`` (:170-211) and the ``do2d``
dispatch (:224-255) for the gate-name forms the reference uses: integer / ``'P#'`` physical gates
(qarray_base_class.py:128-137) and ``'vP#'`` virtual gates with full crosstalk (qarray_base_class.py:143-154).

Layout: the returned grid is ``(ny, nx, n_gate)``, row-major, fast axis = x.

``affine_*`` return the same grid as ``(v0, dx, dy)`` with ``v[iy, ix] = v0 + ix*dx + iy*dy`` -- the form the CUDA
kernel consumes (the grid is never materialised on the device).
"""
from __future__ import annotations

import re

import numpy as np

_P = re.compile(r"^P(\d+)$")
_VP = re.compile(r"^vP(\d+)$")


def _gate_index(gate):
    """Return ('P', idx0) or ('vP', idx0) with 0-based index."""
    if isinstance(gate, (int, np.integer)):
        return "P", int(gate) - 1
    m = _P.match(gate)
    if m:
        return "P", int(m.group(1)) - 1
    m = _VP.match(gate)
    if m:
        return "vP", int(m.group(1)) - 1
    raise ValueError(f"Invalid gate {gate}")


def do2d(n_gate, x_gate, x_min, x_max, x_res, y_gate, y_min, y_max, y_res,
         virtual_gate_matrix=None, virtual_gate_origin=None):
    """``vx[None, :] + vy[:, None]`` of two 1-d sweeps (GateVoltageComposer.py:253-255)."""
    def one(gate, lo, hi, res):
        kind, idx = _gate_index(gate)
        lin = np.linspace(lo, hi, res)
        if kind == "P":
            v = np.zeros((res, n_gate))
            v[:, idx] = lin
            return v
        vd = np.zeros((res, n_gate))
        vd[:, idx] = lin
        return np.einsum("ij,...j->...i", virtual_gate_matrix, vd) + virtual_gate_origin
    vx = one(x_gate, x_min, x_max, x_res)
    vy = one(y_gate, y_min, y_max, y_res)
    return vx[np.newaxis, :] + vy[:, np.newaxis]


def do2d_virtual_coupled(n_gate, dot_x, x_min, x_max, x_res, dot_y, y_min, y_max, y_res, gate_voltages,
                         virtual_gate_matrix, virtual_gate_origin):
    """All dots at ``gate_voltages``; swept dots overridden; ``vg = VGM.Vd + origin`` (GateVoltageComposer.py:170-211).

    ``dot_x`` / ``dot_y`` are 1-based dot numbers (the ``#`` of ``'vP#'``).
    """
    gate_voltages = np.asarray(gate_voltages, dtype=np.float64)
    assert gate_voltages.shape == (n_gate,)
    sweep_x = np.linspace(x_min, x_max, x_res)
    sweep_y = np.linspace(y_min, y_max, y_res)
    vd = np.zeros((x_res, y_res, n_gate))          # reference allocates sizes=[len(x), len(y)] (square scans only)
    vd[:] = gate_voltages
    vd[:, :, dot_x - 1] = sweep_x[np.newaxis, :]
    vd[:, :, dot_y - 1] = sweep_y[:, np.newaxis]
    return np.einsum("ij,...j->...i", virtual_gate_matrix, vd) + virtual_gate_origin


def affine_physical(n_gate, x_gate0, x_min, x_max, x_res, y_gate0, y_min, y_max, y_res):
    """(v0, dx, dy) of a physical-gate scan; gate indices 0-based."""
    v0 = np.zeros(n_gate)
    dx = np.zeros(n_gate)
    dy = np.zeros(n_gate)
    v0[x_gate0] += x_min
    v0[y_gate0] += y_min
    dx[x_gate0] = (x_max - x_min) / (x_res - 1) if x_res > 1 else 0.0
    dy[y_gate0] = (y_max - y_min) / (y_res - 1) if y_res > 1 else 0.0
    return v0, dx, dy


def affine_virtual_coupled(n_gate, dot_x0, x_min, x_max, x_res, dot_y0, y_min, y_max, y_res, gate_voltages,
                           virtual_gate_matrix, virtual_gate_origin):
    """(v0, dx, dy) of a coupled virtual-gate scan; dot indices 0-based."""
    vgm = np.asarray(virtual_gate_matrix, dtype=np.float64)
    base = np.array(gate_voltages, dtype=np.float64, copy=True)
    base[dot_x0] = x_min
    base[dot_y0] = y_min
    v0 = vgm @ base + np.asarray(virtual_gate_origin, dtype=np.float64)
    sx = (x_max - x_min) / (x_res - 1) if x_res > 1 else 0.0
    sy = (y_max - y_min) / (y_res - 1) if y_res > 1 else 0.0
    return v0, vgm[:, dot_x0] * sx, vgm[:, dot_y0] * sy


def affine_grid(v0, dx, dy, nx, ny):
    """Materialise ``v[iy, ix] = v0 + ix*dx + iy*dy`` exactly as the kernel evaluates it (two FMAs per gate)."""
    ix = np.arange(nx, dtype=np.float64)[None, :, None]
    iy = np.arange(ny, dtype=np.float64)[:, None, None]
    return (np.asarray(v0)[None, None, :] + ix * np.asarray(dx)[None, None, :]) + iy * np.asarray(dy)[None, None, :]
