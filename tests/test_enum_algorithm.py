"""CPU property tests of the claims qd_tunnel_select2_kernel (csrc/qd_tunnel_enum.cuh) relies on, against the exhaustive
4^N enumeration of the reference's formulation (charge_states.py:135-222, restated in oracle/path_b.py): no GPU involved.

* E(z) = z^T C z = sum_k d_k y_k^2 on C = L D L^T, y_k depending on dots k..N-1 only (the level structure);
* the level-by-level enumeration with the partial-sum test and the interval lower bound of the levels still to come returns
  exactly the candidates with E <= tau (nothing pruned that belongs, nothing kept that does not);
* with tau = the largest energy of the previous pixel's 32 states re-evaluated at this pixel, the 32 best leaves by
  (energy, index) are the reference's selection.
"""
import itertools
import os
import sys

import numpy as np
import pytest

from oracle import path_b

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _cinv(rng, n):
    off = rng.uniform(0.0, 0.2, size=(n, n))
    off = np.triu(off, 1)
    off = off + off.T
    return np.linalg.inv(np.diag(off.sum(1) + rng.uniform(0.9, 1.3, size=n)) - off)


def _all_candidates(cinv, g, fl):
    n = len(g)
    out = []
    for idx, digs in enumerate(itertools.product((-1, 0, 1, 2), repeat=n)):
        s = fl + np.array(digs)
        if (s < 0).any():
            continue
        z = s - g
        out.append((float(z @ cinv @ z), idx, tuple(int(v) for v in s)))
    return out


@pytest.mark.parametrize("n_dot", [3, 4, 5])
def test_level_sum_is_the_quadratic_form(n_dot):
    import proto_se
    rng = np.random.default_rng(70 + n_dot)
    c = _cinv(rng, n_dot)
    L, d = proto_se.ldl(c)
    np.testing.assert_allclose(L @ np.diag(d) @ L.T, c, atol=1e-14)
    for _ in range(20):
        z = rng.normal(size=n_dot) * 2
        y = L.T @ z
        assert abs((d * y * y).sum() - z @ c @ z) < 1e-12 * (1 + z @ c @ z)
        for k in range(n_dot):                            # y_k sees dots k..N-1 only
            z2 = z.copy()
            z2[:k] += rng.normal(size=k)
            assert abs((L.T @ z2)[k] - y[k]) < 1e-13


@pytest.mark.parametrize("n_dot,empty", [(4, False), (4, True), (5, True), (6, False)])
def test_enumeration_returns_exactly_the_candidates_below_tau(n_dot, empty):
    import proto_se
    rng = np.random.default_rng(80 + n_dot + empty)
    c = _cinv(rng, n_dot)
    for trial in range(6):
        g = rng.uniform(0.2, 4.0, size=n_dot)
        if empty:                                        # deeply empty dots: every digit of theirs costs ~tau
            g[rng.integers(0, n_dot)] = -rng.uniform(1.0, 6.0)
            g[rng.integers(0, n_dot)] = -rng.uniform(0.0, 1.0)
        fl = np.floor(path_b.continuous_ground_state(g[None], c, None)[0])
        cand = _all_candidates(c, g, fl)
        es = sorted(e for e, _, _ in cand)
        for tau in (es[0], es[31], es[min(40, len(es) - 1)] + 1e-9):
            want = {i for e, i, _ in cand if e <= tau * (1 + 1e-12)}
            for use_rem in (False, True):
                counts, nodes = proto_se.enumerate_levels(c, fl - g, fl, tau, np.arange(n_dot), use_rem=use_rem)
                got = set()
                for digs, pe in nodes:
                    idx = 0
                    for j, dg in enumerate(digs):         # digs: dot 0 first (natural order)
                        idx |= (dg + 1) << (2 * (n_dot - 1 - j))
                    got.add(idx)
                assert got == want, (trial, tau, use_rem, len(got), len(want))
            # the bound only removes nodes, never leaves
            c0, _ = proto_se.enumerate_levels(c, fl - g, fl, tau, np.arange(n_dot), use_rem=False)
            c1, _ = proto_se.enumerate_levels(c, fl - g, fl, tau, np.arange(n_dot), use_rem=True)
            assert all(b <= a_ for a_, b in zip(c0, c1)) and c0[-1] == c1[-1]


def test_tau_of_the_previous_basis_selects_the_reference_states():
    import proto_se
    rng = np.random.default_rng(90)
    n_dot = 5
    c = _cinv(rng, n_dot)
    g0 = rng.uniform(0.5, 3.0, size=n_dot)
    g = g0[None, :] + np.arange(24)[:, None] * np.array([0.06, 0.06, 0, 0, 0])[None, :]
    n_c = path_b.continuous_ground_state(g, c, None)
    st = path_b.select_charge_states(g, n_c, c, 32, 1000)
    fl = np.floor(n_c)
    for p in range(1, len(g)):
        prev = st[p - 1].astype(float)
        ok = ((prev - fl[p] >= -1) & (prev - fl[p] <= 2)).all(axis=1)
        if ok.sum() < 32:
            continue
        zz = prev - g[p]
        tau = np.einsum("mi,ij,mj->m", zz, c, zz).max()
        counts, nodes = proto_se.enumerate_levels(c, fl[p] - g[p], fl[p], tau, np.arange(n_dot))
        assert len(nodes) >= 32
        leaves = []
        for digs, pe in nodes:
            idx = 0
            for j, dg in enumerate(digs):
                idx |= (dg + 1) << (2 * (n_dot - 1 - j))
            leaves.append((pe, idx, tuple(int(fl[p][j] + dg) for j, dg in enumerate(digs))))
        best = sorted(leaves)[:32]
        assert {s for _, _, s in best} == {tuple(int(v) for v in s) for s in st[p]}
