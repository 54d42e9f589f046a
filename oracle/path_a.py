"""Path A: constant-interaction open-array ground state.

Test infrastructure (see ``oracle/__init__.py``).  The algorithm is qarray==1.6.0's ``ground_state_open`` (absent from
/root/reference).  PARTLY PINNED: the ``default`` algorithm at T = 0 reproduces, bit for bit, fixtures made by executing
the reference's in-tree mirrors of the upstream functions (free energy + floor/ceil enumeration,
src/qarray_latched/functions.py:30-47) on the exact solution of the reference's relaxation QP (functions.py:66-81) --
``tests/golden/ref_a_*.npz``, ``tests/test_reference_golden.py``.  PARITY UNPINNED for the rest (solver tolerance of the
ADMM/OSQP relaxation, ``thresholded``, ``brute_force``, ``T > 0``).  Restated from its published form and anchored on:

* the reference's call site  src/qadapt/environment/qarray_base_class.py:744-756 (``algorithm``, ``implementation``,
  ``max_charge_carriers``, ``T`` passed through),
* the in-tree mirrors of the upstream JAX code: free energy ``(n - cgd v)^T cdd_inv (n - cgd v)``
  (src/qarray_latched/functions.py:30-34), floor/ceil enumeration (src/qarray_latched/functions.py:36-47), the
  relaxation as the QP  ``min 1/2 n^T cdd_inv n - (cdd_inv cgd v)^T n, n >= 0``
  (src/qarray_latched/functions.py:66-81),
* the algorithm x implementation table  src/qarray_latched/DotArrays/_helper_functions.py:202-210,
* ``k_B`` and ``kT = k_B*T``  src/qarray_latched/DotArrays/_helper_functions.py:213-214, ground_state.py:48.

Decisions:

* matrices are the *dot-only* Maxwell matrices ``cdd_inv (N,N)``, ``cgd (N,G)``  (B.1);
* the relaxation is solved *exactly*.  ``cdd`` is a strictly diagonally dominant M-matrix by construction
  (``convert_to_maxwell``: positive diagonal = row sums, non-positive off-diagonals), so the KKT system of the QP is a
  linear complementarity problem with a K-matrix and is solved by the monotone active-set scheme (Chandrasekaran):
  clamp the dots whose occupation is negative, re-solve, clamp any newly negative ones, stop when none -- at most N
  rounds, no tolerance (B.2).  Upstream uses an ADMM solver (tol ~1e-3) / OSQP here; after ``floor`` the candidate box
  is the same except within that tolerance of an integer occupation;
* ``default``: candidates ``floor(n_c) + {0,1}^N``; ``thresholded``: dot i keeps both iff
  ``abs(frac_i - 1/2) < threshold/2`` else ``round``; ``brute_force``: all ``{0..max}^N`` (B.3, B.6);
* enumeration order: dot 0 slowest, first minimum wins (B.4);
* ``kT > 0``: Boltzmann average over the same candidates (B.5).
"""
from __future__ import annotations

import itertools

import numpy as np

K_B = 8.617333262145e-5  # eV/K  (_helper_functions.py:213-214)


def continuous_relaxation(g: np.ndarray, cdd: np.ndarray) -> np.ndarray:
    """Exact minimiser of ``(n-g)^T cdd^{-1} (n-g)`` over ``n >= 0`` for every row of ``g`` (P, N)."""
    g = np.asarray(g, dtype=np.float64)
    cdd = np.asarray(cdd, dtype=np.float64)
    n_dot = g.shape[1]
    out = g.copy()
    idx = np.nonzero((g < 0).any(axis=1))[0]
    if idx.size == 0:
        return out
    gs = g[idx]
    active = gs < 0
    eye = np.eye(n_dot)
    w = gs
    for _ in range(n_dot + 1):
        both = active[:, :, None] & active[:, None, :]
        m = np.where(both, cdd[None], eye[None])
        rhs = np.where(active, -gs, 0.0)
        mu = np.linalg.solve(m, rhs[..., None])[..., 0]
        w = gs + mu @ cdd.T
        w[active] = 0.0
        new_active = active | (w < 0)
        if np.array_equal(new_active, active):
            break
        active = new_active
    out[idx] = np.maximum(w, 0.0)
    return out


def _binary_deltas(n_dot: int) -> np.ndarray:
    """{0,1}^N, dot 0 slowest (most significant bit)."""
    return np.array(list(itertools.product((0, 1), repeat=n_dot)), dtype=np.float64)


def brute_force_configurations(n_dot: int, max_charge_carriers: int, sum_filter: bool = False) -> np.ndarray:
    confs = np.array(list(itertools.product(range(max_charge_carriers + 1), repeat=n_dot)), dtype=np.float64)
    if sum_filter:
        confs = confs[confs.sum(axis=1) <= max_charge_carriers]
    return confs


def _select(energies: np.ndarray, confs: np.ndarray, kT: float):
    """energies (p, M) [inf = excluded], confs (p, M, N) or (M, N) -> (n (p, N), margin (p,))."""
    best = np.argmin(energies, axis=1)
    e_best = np.take_along_axis(energies, best[:, None], axis=1)
    if energies.shape[1] > 1:
        part = np.partition(energies, 1, axis=1)
        margin = part[:, 1] - part[:, 0]
    else:
        margin = np.full(energies.shape[0], np.inf)
    if kT > 0.0:
        w = np.exp(-(energies - e_best) / kT)
        if confs.ndim == 2:
            n = (w @ confs) / w.sum(axis=1, keepdims=True)
        else:
            n = np.einsum("pm,pmd->pd", w, confs) / w.sum(axis=1, keepdims=True)
    else:
        if confs.ndim == 2:
            n = confs[best]
        else:
            n = np.take_along_axis(confs, best[:, None, None], axis=1)[:, 0, :]
    return n, margin


def ground_state_open(vg, cgd, cdd_inv, cdd, algorithm: str = "default", threshold: float = 1.0,
                      max_charge_carriers: int = 4, kT: float = 0.0, brute_sum_filter: bool = False,
                      chunk: int = 2048, return_margin: bool = False):
    """Ground-state occupations for voltages ``vg`` (P, G) -> (P, N) float64 (integers when ``kT == 0``)."""
    vg = np.asarray(vg, dtype=np.float64)
    cgd = np.asarray(cgd, dtype=np.float64)
    cdd_inv = np.asarray(cdd_inv, dtype=np.float64)
    n_dot = cdd_inv.shape[0]
    if vg.shape[-1] != cgd.shape[1]:
        raise ValueError(f"The shape of vg is in correct it should be of shape (..., n_gate) = (...,{cgd.shape[1]})")
    algorithm = algorithm.lower()
    p_total = vg.shape[0]
    n_out = np.empty((p_total, n_dot))
    margin_out = np.empty(p_total)
    deltas = _binary_deltas(n_dot)
    brute = brute_force_configurations(n_dot, max_charge_carriers, brute_sum_filter) if algorithm == "brute_force" else None
    for s in range(0, p_total, chunk):
        v = vg[s:s + chunk]
        g = v @ cgd.T                                      # (p, N)   v_dash = cgd @ vg
        if algorithm == "brute_force":
            r = brute[None, :, :] - g[:, None, :]
            e = np.einsum("pmi,ij,pmj->pm", r, cdd_inv, r)
            n, margin = _select(e, brute, kT)
        elif algorithm in ("default", "thresholded"):
            n_c = continuous_relaxation(g, cdd)
            fl = np.floor(n_c)
            confs = fl[:, None, :] + deltas[None, :, :]    # (p, 2^N, N)
            r = confs - g[:, None, :]
            e = np.einsum("pmi,ij,pmj->pm", r, cdd_inv, r)
            if algorithm == "thresholded":
                frac = n_c - fl
                both = np.abs(frac - 0.5) < threshold / 2.0
                rounded = np.floor(n_c + 0.5) - fl           # 0 or 1: the delta kept when only one is kept (half away from 0)
                ok = both[:, None, :] | (deltas[None, :, :] == rounded[:, None, :])
                e = np.where(ok.all(axis=2), e, np.inf)
            n, margin = _select(e, confs, kT)
        else:
            raise AssertionError(f"Algorithm {algorithm} not supported")
        n_out[s:s + chunk] = n
        margin_out[s:s + chunk] = margin
    if return_margin:
        return n_out, margin_out
    return n_out
