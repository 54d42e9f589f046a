"""GateVoltageComposer with the reference's interface (src/qarray_latched/DotArrays/GateVoltageComposer.py:16-282).

Every scan the reference builds is affine in the two pixel indices, ``v[iy, ix] = v0 + ix*dx + iy*dy``.  The class
therefore has two faces: ``do1d`` / ``do2d`` / ``meshgrid*`` return the materialised ndarray exactly like the reference
(for callers that want the grid), and ``affine2d`` returns ``(v0, dx, dy)`` -- the form the CUDA kernel consumes, so the
hot path never touches an O(pixels x gates) array.

Gate names: ``int`` or ``'P#'`` physical gate, ``'vP#'`` virtual plunger, ``'e#_#'`` detuning (difference of two virtual
sweeps), ``'U#_#'`` on-site (their sum / sqrt 2); numbering is 1-based.
"""
from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np

_PATTERNS = (("P", re.compile(r"^P(\d+)$")), ("vP", re.compile(r"^vP(\d+)$")),
             ("e", re.compile(r"^e(\d+)_(\d+)$")), ("U", re.compile(r"^U(\d+)_(\d+)$")))


@dataclass
class GateVoltageComposer:
    n_gate: int
    n_dot: int | None = None
    n_sensor: int | None = 0
    virtual_gate_origin: np.ndarray | None = None
    virtual_gate_matrix: np.ndarray | None = None

    # ---- validation ------------------------------------------------------------------------------------------
    def _check_gate(self, gate):
        assert isinstance(gate, (int, np.integer)), "gate must be an int"
        assert 1 <= gate <= self.n_gate, f"gate must be in the range 1 to {self.n_gate}"

    def _check_dot(self, dot):
        assert isinstance(dot, (int, np.integer)), "dot must be an int"
        assert 1 <= dot <= self.n_dot, f"dot must be in the range 1 to {self.n_dot}"

    def _check_virtual(self):
        assert self.virtual_gate_origin is not None, "virtual_gate_origin must be set"
        assert self.virtual_gate_matrix is not None, "virtual_gate_matrix must be set"
        assert self.n_dot is not None, "n_dot must be set"

    # ---- direction of a named sweep: v(t) = base + t * direction, t the swept value -----------------------------
    def _direction(self, gate):
        """Returns (direction (n_gate,), offset (n_gate,)) such that a sweep of ``gate`` over values t is
        ``offset + t * direction``."""
        zero = np.zeros(self.n_gate)
        if isinstance(gate, (int, np.integer)):
            self._check_gate(gate)
            d = zero.copy()
            d[gate - 1] = 1.0
            return d, zero
        if not isinstance(gate, str):
            raise ValueError(f"Invalid gate {gate}")
        for kind, pat in _PATTERNS:
            m = pat.match(gate)
            if not m:
                continue
            idx = [int(g) for g in m.groups()]
            if kind == "P":
                self._check_gate(idx[0])
                d = zero.copy()
                d[idx[0] - 1] = 1.0
                return d, zero
            self._check_virtual()
            vgm = np.asarray(self.virtual_gate_matrix, dtype=np.float64)
            origin = np.asarray(self.virtual_gate_origin, dtype=np.float64)
            for i in idx:
                self._check_dot(i)
            if kind == "vP":
                return vgm[:, idx[0] - 1].copy(), origin
            if kind == "e":       # (VGM e_a t + o) - (VGM e_b t + o)
                return vgm[:, idx[0] - 1] - vgm[:, idx[1] - 1], zero
            return (vgm[:, idx[0] - 1] + vgm[:, idx[1] - 1]) / np.sqrt(2), 2 * origin / np.sqrt(2)
        raise ValueError(f"Invalid gate {gate} must be in the form P[int], vP[int], e[int]_[int], U[int]_[int]")

    # ---- affine descriptors (what the kernel takes) -------------------------------------------------------------
    def affine2d(self, x_gate, x_min, x_max, x_res, y_gate, y_min, y_max, y_res, gate_voltages=None,
                 add_full_crosstalk: bool = False):
        """(v0, dx, dy) with ``v[iy, ix] = v0 + ix*dx + iy*dy`` == ``do2d(...)[iy, ix]``."""
        sx = (x_max - x_min) / (x_res - 1) if x_res > 1 else 0.0
        sy = (y_max - y_min) / (y_res - 1) if y_res > 1 else 0.0
        if add_full_crosstalk:
            assert gate_voltages is not None, "gate_voltages must be provided to add full crosstalk"
            dots = []
            for gate in (x_gate, y_gate):
                m = _PATTERNS[1][1].match(gate) if isinstance(gate, str) else None
                if not m:
                    raise ValueError(f"Gate {gate} must be a virtual gate in the form vP[int] when using add_full_crosstalk")
                dots.append(int(m.group(1)))
            self._check_virtual()
            gv = np.array(gate_voltages, dtype=np.float64)
            assert gv.shape == (self.n_dot + 1,), "gate voltages do not match the number of dots (including the sensor dot)"
            for d in dots:
                self._check_dot(d)
            vgm = np.asarray(self.virtual_gate_matrix, dtype=np.float64)
            base = gv.copy()
            base[dots[0] - 1] = x_min
            base[dots[1] - 1] = y_min
            v0 = vgm @ base + np.asarray(self.virtual_gate_origin, dtype=np.float64)
            return v0, vgm[:, dots[0] - 1] * sx, vgm[:, dots[1] - 1] * sy
        dxv, ox = self._direction(x_gate)
        dyv, oy = self._direction(y_gate)
        return ox + oy + x_min * dxv + y_min * dyv, dxv * sx, dyv * sy

    # ---- materialised grids (reference return conventions) -----------------------------------------------------
    def do1d(self, gate, min, max, res):  # noqa: A002  (reference argument names)
        d, o = self._direction(gate)
        return o[None, :] + np.linspace(min, max, res)[:, None] * d[None, :]

    def do2d(self, x_gate, x_min, x_max, x_res, y_gate, y_min, y_max, y_res, gate_voltages=None,
             add_full_crosstalk: bool = False):
        if add_full_crosstalk:
            if isinstance(gate_voltages, np.ndarray):
                gate_voltages = gate_voltages.tolist()
            v0, dx, dy = self.affine2d(x_gate, x_min, x_max, x_res, y_gate, y_min, y_max, y_res, gate_voltages, True)
            if x_res != y_res:
                raise ValueError("coupled virtual scans are square in the reference (meshgrid_virtual_coupled)")
            vgm = np.asarray(self.virtual_gate_matrix, dtype=np.float64)
            vd = np.empty((y_res, x_res, self.n_gate))
            vd[:] = np.asarray(gate_voltages, dtype=np.float64)
            dx_dot = int(_PATTERNS[1][1].match(x_gate).group(1)) - 1
            dy_dot = int(_PATTERNS[1][1].match(y_gate).group(1)) - 1
            vd[:, :, dx_dot] = np.linspace(x_min, x_max, x_res)[None, :]
            vd[:, :, dy_dot] = np.linspace(y_min, y_max, y_res)[:, None]
            return vd @ vgm.T + np.asarray(self.virtual_gate_origin, dtype=np.float64)
        return self.do1d(x_gate, x_min, x_max, x_res)[None, :, :] + self.do1d(y_gate, y_min, y_max, y_res)[:, None, :]

    def meshgrid(self, gates, arrays):
        assert all(np.ndim(a) == 1 for a in arrays), "arrays must be 1d"
        assert len(gates) == len(arrays), "gates and arrays must be the same length"
        for g in gates:
            self._check_gate(g)
        grids = np.meshgrid(*arrays)
        out = np.zeros(grids[0].shape + (self.n_gate,)) if grids else np.zeros((self.n_gate,))
        for g, grid in zip(gates, grids):
            out[..., g - 1] = grid
        return out

    def meshgrid_virtual(self, dots, arrays):
        self._check_virtual()
        assert all(np.ndim(a) == 1 for a in arrays), "arrays must be 1d"
        assert len(dots) == len(arrays), "gates and arrays must be the same length"
        for d in dots:
            self._check_dot(d)
        grids = np.meshgrid(*arrays)
        vd = np.zeros(grids[0].shape + (self.n_dot + self.n_sensor,))
        for d, grid in zip(dots, grids):
            vd[..., d - 1] = grid
        return vd @ np.asarray(self.virtual_gate_matrix).T + self.virtual_gate_origin

    def meshgrid_virtual_coupled(self, dots, arrays, gate_voltages):
        if len(dots) != 2:
            raise NotImplementedError("meshgrid_virtual_coupled currently only supports a two-dot sweep")
        x, y = np.asarray(arrays[0]), np.asarray(arrays[1])
        return self.do2d(f"vP{dots[0]}", x[0], x[-1], x.size, f"vP{dots[1]}", y[0], y[-1], y.size, gate_voltages, True)
