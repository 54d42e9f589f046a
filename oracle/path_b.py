"""Path B: the tunnel-coupled ground state that QADAPT's ``env.step`` actually runs (SURVEY.md section 8a rows B1-B7,
Appendix C).

Test infrastructure (see ``oracle/__init__.py``).  Unlike Path A this code IS in the reference tree; the restatement is
literal, in NumPy fp64, following

* src/qarray_latched/DotArrays/ground_state.py:24-166          (``_ground_state_open``: orchestration)
* src/qarray_latched/DotArrays/charge_states.py:36-88           (continuous relaxation: closed form or 50 projected
                                                                  gradient steps, lr 0.1)
* src/qarray_latched/DotArrays/charge_states.py:135-222         (4^N candidates floor + {-1,0,1,2}^N in base-4 index
                                                                  order, last dot fastest; negative -> +inf; per-1000
                                                                  chunk stable top-32 merged with the running best)
* src/qarray_latched/DotArrays/hamiltonian_build.py:12-45       (free energy of the kept states, recomputed unmasked)
* src/qarray_latched/DotArrays/hamiltonian_build.py:75-137      (nearest-neighbour tunnelling, ``fermionic_negative``)
* src/qarray_latched/DotArrays/hamiltonian_build.py:460-483     (diag(F))
* src/qarray_latched/DotArrays/voltage_dependent_capacitance.py:78-91, 128-141 and ground_state.py:53-58 (optional
                                                                  linear model: per pixel ``cdd = cdd_0 (1 + alpha mean|v|)``,
                                                                  ``cgd = cgd_0 (1 + beta mean|v|)`` over ALL entries of v_ext;
                                                                  in the ground state only -- the sensor keeps the constant
                                                                  matrices, TunnelCoupledChargeSensed.py:342-376)
* src/qarray_latched/DotArrays/barrier_voltage_model.py:55-151  (``vb_eff = vb + Cbg vg``; the cross-barrier term is the
                                                                  diagonal of a zero-diagonal matrix = 0;
                                                                  ``t = tc_base exp(-alpha vb_eff)``, no abs)

PINNED: ``tests/golden/ref_*tunnel*.npz`` hold the output of the reference's own code (the files above, executed unmodified
on a NumPy stand-in for jax, ``tests/golden/refshim.py``); this restatement reproduces them to 5e-14
(``tests/test_reference_golden.py``).
``T`` is read by the reference and never used on this path (ground_state.py:48).
"""
from __future__ import annotations

import numpy as np


def continuous_ground_state(g: np.ndarray, cinv: np.ndarray, scale=None) -> np.ndarray:
    """charge_states.py:36-88 -- ``g`` (P, N) = cgd[:N] @ v_ext; ``scale`` (P,): per-pixel factor of cdd (cdd_inv is
    divided by it), None = 1."""
    n_c = g.copy()
    bad = (g < 0).any(axis=1)
    if bad.any():
        gb = g[bad]
        n = np.clip(gb, 0, None)
        sc = 1.0 if scale is None else np.asarray(scale)[bad][:, None]
        cg = (gb @ cinv.T) / sc                        # cdd_inv @ (cgd @ v)
        for _ in range(50):
            grad = (n @ cinv.T) / sc - cg
            n = np.clip(n - 0.1 * grad, 0, None)
        n_c[bad] = n
    return np.clip(n_c, 0, None)


def select_charge_states(g, n_c, cinv, num_states: int = 32, chunk_size: int = 1000):
    """charge_states.py:135-222 -- returns int states (P, num_states, N)."""
    p, n_dot = g.shape
    floor_values = np.floor(n_c)
    total = 4 ** n_dot
    if not chunk_size:
        chunk_size = total
    n_chunks = (total + chunk_size - 1) // chunk_size
    best_e = np.full((p, num_states), np.inf)
    best_s = np.zeros((p, num_states, n_dot))
    for c in range(n_chunks):
        idx = np.arange(chunk_size) + c * chunk_size
        within = idx < total
        safe = idx % total
        digits = np.zeros((chunk_size, n_dot), dtype=np.int64)
        tmp = safe.copy()
        for i in range(n_dot):
            digits[:, n_dot - 1 - i] = tmp % 4
            tmp //= 4
        deltas = np.array([-1, 0, 1, 2])[digits]                       # (chunk, N)
        confs = deltas[None, :, :] + floor_values[:, None, :]           # (P, chunk, N)
        valid = (confs >= 0).all(axis=-1) & within[None, :]
        r = confs - g[:, None, :]
        e = np.einsum("pci,ij,pcj->pc", r, cinv, r)
        e = np.where(valid, e, np.inf)
        order = np.argsort(e, axis=1, kind="stable")[:, :num_states]
        ce = np.take_along_axis(e, order, axis=1)
        cs = np.take_along_axis(confs, order[:, :, None], axis=1)
        if ce.shape[1] < num_states:                                    # chunk smaller than the basis (never in practice)
            padn = num_states - ce.shape[1]
            ce = np.concatenate([ce, np.full((p, padn), np.inf)], axis=1)
            cs = np.concatenate([cs, np.zeros((p, padn, n_dot))], axis=1)
        comb_e = np.concatenate([best_e, ce], axis=1)
        comb_s = np.concatenate([best_s, cs], axis=1)
        fin = np.argsort(comb_e, axis=1, kind="stable")[:, :num_states]
        best_e = np.take_along_axis(comb_e, fin, axis=1)
        best_s = np.take_along_axis(comb_s, fin[:, :, None], axis=1)
    return best_s.astype(np.int64)


def tunnel_couplings(m, v_ext):
    """(P, N-1) nearest-neighbour couplings."""
    n_dot = m.cdd_inv.shape[0]
    p = v_ext.shape[0]
    n_gate = m.n_gate
    if m.cbg is not None and v_ext.shape[1] > n_gate:
        vg, vb = v_ext[:, :n_gate], v_ext[:, n_gate:]
        vb_eff = vb + vg @ np.asarray(m.cbg).T
        alpha = np.asarray(m.alpha, dtype=np.float64)[:n_dot - 1]
        return m.tc_base * np.exp(-alpha[None, :] * vb_eff[:, :n_dot - 1])
    return np.full((p, n_dot - 1), float(m.tc_base))


def hamiltonian(states, g, cinv, t):
    """states (P, M, N) int, g (P, N), t (P, N-1) -> H (P, M, M)."""
    p, mm, n_dot = states.shape
    s = states.astype(np.float64)
    r = s - g[:, None, :]
    f = np.einsum("pmi,ij,pmj->pm", r, cinv, r)
    h = np.zeros((p, mm, mm))
    idx = np.arange(mm)
    h[:, idx, idx] = f
    si = s[:, :, None, :]
    sj = s[:, None, :, :]
    diff = sj - si
    for d in range(n_dot - 1):
        exp = np.zeros(n_dot)
        exp[d], exp[d + 1] = -1, 1
        fwd = (diff == exp).all(axis=-1)
        bwd = (diff == -exp).all(axis=-1)
        n_from = si[..., d]
        n_to = si[..., d + 1]
        with np.errstate(invalid="ignore"):
            ef = -t[:, d, None, None] * np.sqrt(n_from * (n_to + 1))
            eb = -t[:, d, None, None] * np.sqrt(n_to * (n_from + 1))
        h = h + fwd * ef + bwd * eb
    return h, f


def ground_state_open(m, v_ext, return_gap: bool = False, chunk: int = 128):
    """``m``: oracle.scan.Model with algorithm == "tunnel" (cdd_inv = cdd_inv_full[:N,:N], cgd = cgd_full[:N]).
    Returns <n> (P, N) (non-integer); optionally the spectral gap lambda_1 - lambda_0 of each pixel's Hamiltonian."""
    v_ext = np.asarray(v_ext, dtype=np.float64)
    cinv = np.asarray(m.cdd_inv, dtype=np.float64)
    a = np.asarray(m.cgd, dtype=np.float64)
    n_dot = cinv.shape[0]
    out = np.empty((v_ext.shape[0], n_dot))
    gap = np.empty(v_ext.shape[0])
    for s0 in range(0, v_ext.shape[0], chunk):
        v = v_ext[s0:s0 + chunk]
        g = v @ a.T
        s_c = None
        if getattr(m, "vc_alpha", 0.0) or getattr(m, "vc_beta", 0.0):
            vmean = np.abs(v).mean(axis=1)
            kind = int(getattr(m, "vc_kind", 0))
            if kind == 1:      # quadratic_voltage_dependent_cdd, voltage_dependent_capacitance.py:94-99
                s_c = 1.0 + m.vc_alpha * (v ** 2).sum(axis=1)
            elif kind == 2:    # sigmoid_voltage_dependent_cdd, :102-109 (jax.nn.sigmoid(x) = 1 / (1 + exp(-x)))
                s_c = 1.0 + m.vc_alpha / (1.0 + np.exp(1.0 - np.linalg.norm(v, axis=1) / m.vc_vchar))
            else:              # linear_voltage_dependent_cdd, :78-83
                s_c = 1.0 + m.vc_alpha * vmean
            g = g * (1.0 + m.vc_beta * vmean)[:, None]
        n_c = continuous_ground_state(g, cinv, s_c)
        states = select_charge_states(g, n_c, cinv, m.num_charge_states, m.charge_state_batch_size)   # order is scale-free
        t = tunnel_couplings(m, v)
        h, f = hamiltonian(states, g, cinv, t)
        if s_c is not None:                              # diag(F) / s_c: remove F, add it back scaled
            idx = np.arange(h.shape[1])
            h[:, idx, idx] += f / s_c[:, None] - f
        w, vec = np.linalg.eigh(h)
        psi2 = np.abs(vec[:, :, 0]) ** 2
        out[s0:s0 + chunk] = np.einsum("pm,pmd->pd", psi2, states.astype(np.float64))
        gap[s0:s0 + chunk] = w[:, 1] - w[:, 0]
    if return_gap:
        return out, gap
    return out
