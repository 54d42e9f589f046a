"""Known-answer and self-consistency tests that pin the CPU oracle (SURVEY.md section 8c lists them; the reference ships
no test for this path, so these are the pins).  CPU only."""
import itertools

import numpy as np
import pytest

from oracle import capacitance as cap
from oracle import composer, latching, noise, path_a, philox, sensor
from oracle import scan as oscan

# literal 2-dot matrices of the reference's own demo (src/qarray_latched/DotArrays/ground_state.py:190-205)
CDD = np.array([[0, .4], [.4, 0]])
CGD = np.array([[1, .4, 0], [.4, 1, 0]])
CDS = np.array([[.05, .04]])
CGS = np.array([[.06, .05, 1]])


def _random_device(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.uniform(0, 0.2, (n, n))
    a = np.triu(a, 1)
    a = a + a.T
    g = np.zeros((n, n + 1))
    g[:, :n] = rng.uniform(0.3, 1.0, (n, n))
    return a, g, rng.uniform(0.03, 0.05, (1, n)), np.concatenate([rng.uniform(0, 1e-4, (1, n)), [[0.97]]], axis=1)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    f = lambda t: tuple(int(x) for x in t)  # noqa: E731
    assert f(philox.philox4x32_10(0, 0, 0, 0, 0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    m = 0xffffffff
    assert f(philox.philox4x32_10(m, m, m, m, m, m)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert f(philox.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_philox_draw_statistics():
    d = philox.pixel_draws(12345, 200_000)
    for k in ("z_white", "z_radial"):
        assert abs(d[k].mean()) < 0.01 and abs(d[k].std() - 1) < 0.01
    for k in ("u_latch", "u_tele"):
        assert abs(d[k].mean() - 0.5) < 0.005 and d[k].min() >= 0 and d[k].max() < 1
    assert abs(np.corrcoef(d["z_white"], d["z_radial"])[0, 1]) < 0.01


def test_maxwell_conversion_kat():
    cdd_f, cdi_f, cgd_f = cap.with_sensor(CDD, CGD, CDS, CGS)
    assert np.allclose(cdd_f @ cdi_f, np.eye(3), atol=1e-14)
    full_nm = np.block([[CDD, CDS.T], [CDS, np.zeros((1, 1))]])
    assert np.allclose(np.diag(cdd_f), full_nm.sum(1) + np.vstack([CGD, CGS]).sum(1))      # row-sum rule
    assert np.allclose(cdd_f - np.diag(np.diag(cdd_f)), -full_nm)
    assert np.array_equal(cgd_f, -np.vstack([CGD, CGS]))
    cdd, cdi, cgd = cap.convert_to_maxwell(CDD, CGD)
    assert np.allclose(cdd, [[1.8, -.4], [-.4, 1.8]])
    # barriers enter as extra voltage columns only
    cbd = np.array([[.06, ], [.05, ]])
    cbs = np.array([[.001]])
    cdd_b, cdi_b, cgd_b = cap.with_barriers_and_sensor(CDD, CGD, CDS, CGS, cbd, cbs)
    assert cgd_b.shape == (3, 4) and np.array_equal(cgd_b[:, 3], -np.array([.06, .05, .001]))
    assert np.allclose(np.diag(cdd_b), np.diag(cdd_f) + np.array([.06, .05, .001]))


@pytest.mark.parametrize("n", [2, 3, 5])
def test_relaxation_is_the_exact_qp_minimiser(n):
    """Monotone active-set solution == brute force over all 2^N KKT active sets."""
    a, g_nm, _, _ = _random_device(n, n)
    cdd, cdi, _ = cap.convert_to_maxwell(a, g_nm)
    rng = np.random.default_rng(0)
    g = rng.normal(0.3, 1.5, (300, n))
    nc = path_a.continuous_relaxation(g, cdd)
    for row, sol in zip(g, nc):
        best = None
        for s in itertools.product([False, True], repeat=n):
            s = np.array(s)
            f = ~s
            cand = np.zeros(n)
            if f.any():
                cand[f] = row[f] + (np.linalg.solve(cdi[np.ix_(f, f)], cdi[np.ix_(f, s)] @ row[s]) if s.any() else 0)
            if (cand < -1e-12).any():
                continue
            e = (cand - row) @ cdi @ (cand - row)
            if best is None or e < best[0] - 1e-13:
                best = (e, cand)
        assert np.allclose(sol, best[1], atol=1e-12)
    assert (nc >= 0).all()


def test_uncoupled_dots_round_independently():
    """Cdd = 0 -> separable energy -> n = round(max(cgd.v, 0)) dot by dot."""
    n = 4
    cgd_nm = np.zeros((n, n + 1))
    cgd_nm[:, :n] = np.diag([1.0, 0.9, 1.1, 0.95])
    cdd, cdi, cgd = cap.convert_to_maxwell(np.zeros((n, n)), cgd_nm)
    rng = np.random.default_rng(1)
    vg = rng.uniform(-4, 1, (2000, n + 1))
    got = path_a.ground_state_open(vg, cgd, cdi, cdd)
    g = vg @ cgd.T
    want = np.floor(np.maximum(g, 0) + 0.5)
    ties = np.abs(np.maximum(g, 0) % 1 - 0.5) < 1e-9
    assert np.array_equal(got[~ties.any(1)], want[~ties.any(1)])


def test_default_equals_brute_force_and_thresholded_one():
    cdd, cdi, cgd = cap.convert_to_maxwell(CDD * 0.3, CGD)          # weak coupling: the floor/ceil box suffices
    v0, dx, dy = composer.affine_physical(3, 0, -3.7, 0.3, 48, 1, -3.9, 0.2, 48)
    vg = composer.affine_grid(v0, dx, dy, 48, 48).reshape(-1, 3) + 1e-3 * np.pi
    d, m = path_a.ground_state_open(vg, cgd, cdi, cdd, "default", return_margin=True)
    b = path_a.ground_state_open(vg, cgd, cdi, cdd, "brute_force", max_charge_carriers=5)
    t = path_a.ground_state_open(vg, cgd, cdi, cdd, "thresholded", threshold=1.0)
    ok = m > 1e-9
    assert ok.mean() > 0.99
    assert np.array_equal(d[ok], b[ok]) and np.array_equal(d[ok], t[ok])
    assert d.max() == 5 and d.min() == 0
    # a tight threshold prunes candidates: never lower in energy than the default answer
    cdd, cdi, cgd = cap.convert_to_maxwell(CDD * 2.0, CGD)          # strong coupling: rounding alone goes wrong
    d = path_a.ground_state_open(vg, cgd, cdi, cdd, "default")
    t2 = path_a.ground_state_open(vg, cgd, cdi, cdd, "thresholded", threshold=0.05)
    e = lambda nn: np.einsum("pi,ij,pj->p", nn - vg @ cgd.T, cdi, nn - vg @ cgd.T)  # noqa: E731
    assert (e(t2) >= e(d) - 1e-12).all() and (t2 != d).any()


def test_thermal_average_limits():
    cdd, cdi, cgd = cap.convert_to_maxwell(CDD, CGD)
    rng = np.random.default_rng(2)
    vg = rng.uniform(-3, 0, (500, 3))
    hard, m = path_a.ground_state_open(vg, cgd, cdi, cdd, return_margin=True)
    cold = path_a.ground_state_open(vg, cgd, cdi, cdd, kT=1e-5)
    assert np.allclose(cold[m > 1e-2], hard[m > 1e-2], atol=1e-9)
    warm = path_a.ground_state_open(vg, cgd, cdi, cdd, kT=0.05)
    assert (np.abs(warm - np.rint(warm)) > 1e-3).any()
    assert (warm >= np.floor(np.minimum(hard, warm)) - 1e-12).all()


def test_optimal_vg_round_trip():
    """ground_state_open(optimal_Vg(n)) == n[:N] (qarray_config.yaml:122 centre: dots 1, sensor 0.53)."""
    for n in (2, 4, 6):
        a, g_nm, cds, cgs = _random_device(n, 10 + n)
        cdd, cdi, cgd = cap.convert_to_maxwell(a, g_nm)
        _, cdi_f, cgd_f = cap.with_sensor(a, g_nm, cds, cgs)
        target = np.array([1.0] * n + [0.53])
        vg = cap.optimal_vg(cdi_f, cgd_f, target)
        assert np.allclose(cgd_f @ vg, target, atol=1e-9)
        got = path_a.ground_state_open(vg[None, :], cgd, cdi, cdd)
        assert np.array_equal(got[0], np.ones(n))


def test_virtual_gate_matrix_sign_conventions():
    _, cdi_f, cgd_f = cap.with_sensor(CDD, CGD, CDS, CGS)
    vgm = cap.optimal_virtual_gate_matrix(cdi_f, cgd_f, "h")
    assert np.allclose(cdi_f @ cgd_f @ vgm, -np.eye(3), atol=1e-12)
    assert np.allclose(cap.optimal_virtual_gate_matrix(cdi_f, cgd_f, "electrons"), -vgm)


def test_composer_grids_and_affine_forms_agree():
    vgm = np.array([[-1.0, 0.2, 0.0], [0.1, -1.1, 0.05], [0.0, 0.0, -1.0]])
    origin = np.array([0.1, -0.2, 0.05])
    phys = composer.do2d(3, 1, -2.0, 1.0, 17, "P2", -1.0, 3.0, 9)
    assert phys.shape == (9, 17, 3) and phys[3, 5, 0] == np.linspace(-2, 1, 17)[5] and phys[3, 5, 1] == np.linspace(-1, 3, 9)[3]
    v0, dx, dy = composer.affine_physical(3, 0, -2.0, 1.0, 17, 1, -1.0, 3.0, 9)
    assert np.allclose(composer.affine_grid(v0, dx, dy, 17, 9), phys, atol=1e-13)
    gv = [0.3, -0.4, 0.7]
    virt = composer.do2d_virtual_coupled(3, 1, -2.0, 1.0, 16, 2, -1.0, 3.0, 16, gv, vgm, origin)
    v0, dx, dy = composer.affine_virtual_coupled(3, 0, -2.0, 1.0, 16, 1, -1.0, 3.0, 16, gv, vgm, origin)
    assert np.allclose(composer.affine_grid(v0, dx, dy, 16, 16), virt, atol=1e-13)
    # fast axis = x = first (left) dot
    assert np.allclose(virt[0, 1] - virt[0, 0], vgm[:, 0] * 3.0 / 15)
    with pytest.raises(ValueError):
        composer.do2d(3, "q1", 0, 1, 4, 1, 0, 1, 4)


def test_sensor_range_and_integer_shift_invariance():
    _, cdi_f, cgd_f = cap.with_sensor(CDD, CGD, CDS, CGS)
    rng = np.random.default_rng(3)
    n = rng.integers(0, 4, (400, 2)).astype(float)
    v = rng.uniform(-3, 0, (400, 3))
    z = sensor.charge_sensor_signal(n, v, cdi_f, cgd_f, 0.1)
    assert z.shape == (400, 1) and (z > 0).all() and (z <= 10).all()
    # moving the sensor gate by exactly one sensor charge re-centres N_sensor: same signal
    shift = np.zeros(3)
    shift[2] = 1.0 / cgd_f[2, 2]
    v2 = v + shift
    n_dash = np.einsum("ij,pj->pi", cgd_f, v2 - v)[:, :2]          # dots feel the shift too: compensate
    z2 = sensor.charge_sensor_signal(n + n_dash, v2, cdi_f, cgd_f, 0.1)
    assert np.allclose(z, z2, rtol=1e-9)


def test_latching_limits_and_acceptance_rate():
    rng = np.random.default_rng(4)
    ny, nx, n = 40, 64, 3
    base = np.cumsum(rng.random((ny, nx, n)) < 0.1, axis=1).astype(float)
    u = rng.random((ny, nx))
    ones, zeros = np.ones(n), np.zeros(n)
    assert np.array_equal(latching.add_latching(base, u, ones, np.ones((n, n))), base)
    frozen = latching.add_latching(base, u, zeros, np.zeros((n, n)))
    single = (np.abs(np.diff(base, axis=1)).sum(-1) <= 2).all()
    if single:
        assert (frozen == frozen[:, :1]).all()
    # carry_rows: a flat pass holds the state across the row end
    flat = latching.add_latching(base, u, zeros, np.zeros((n, n)), carry_rows=True)
    assert (flat == base[0, 0]).all() or not single
    # acceptance frequency of single-dot transitions -> p_leads (binomial 4 sigma)
    steps = np.zeros((2000, 2, 1))
    steps[:, 1, 0] = 1.0
    uu = np.random.default_rng(5).random((2000, 2))
    out = latching.add_latching(steps, uu, np.array([0.3]), np.zeros((1, 1)))
    rate = (out[:, 1, 0] == 1).mean()
    assert abs(rate - 0.3) < 4 * np.sqrt(0.3 * 0.7 / 2000)


def test_telegraph_chain_statistics():
    rng = np.random.default_rng(6)
    p01, p10 = 0.02, 0.05
    u = rng.random((1, 400_000))
    s = noise.telegraph_states(u, np.array([0.9]), p01, p10, carry_rows=True)[0]
    assert abs(s.mean() - p01 / (p01 + p10)) < 0.01
    flips = np.flatnonzero(np.diff(s) != 0)
    runs = np.diff(flips)
    on = runs[::2] if s[flips[0] + 1] == 1 else runs[1::2]
    off = runs[1::2] if s[flips[0] + 1] == 1 else runs[::2]
    assert abs(on.mean() - 1 / p10) / (1 / p10) < 0.05 and abs(off.mean() - 1 / p01) / (1 / p01) < 0.05
    # per-row chains start from the stationary distribution
    u_row = rng.random(4000)
    rows = noise.telegraph_states(rng.random((4000, 1)), u_row, 0.0, 0.0, carry_rows=False)
    assert (rows == 0).all()
    rows = noise.telegraph_states(rng.random((4000, 1)), u_row, 1e-9, 3e-9, carry_rows=False)
    assert abs(rows.mean() - 0.25) < 0.03


def test_radial_noise_modes():
    z = np.ones((8, 8))
    zr = np.arange(64, dtype=float)
    assert np.array_equal(noise.radial_noise(z, zr, 0, 0, 1, 0, 1, 1, 0, 1), z)
    assert np.array_equal(noise.radial_noise(z, zr, 2, 0, 1, 0, 1, 1, 0, 1), zr.reshape(8, 8))
    out = noise.radial_noise(z, zr, 1, 0.0, 1.0, 0.0, 1.0, 0.01, 2.0, 0.05)
    assert out[0, 0] == 1.0 and out[0, 2] == 1.0 and out[7, 7] == 1 + 63 * min(0.05, 0.01 * (np.hypot(7, 7) - 2))


def test_whole_scan_oracle_runs_and_is_deterministic():
    cdd, cdi, cgd = cap.convert_to_maxwell(CDD, CGD)
    _, cdi_f, cgd_f = cap.with_sensor(CDD, CGD, CDS, CGS)
    m = oscan.Model(cdd_inv=cdi, cdd=cdd, cgd=cgd, cdd_inv_full=cdi_f, cgd_full=cgd_f, latching=True,
                    p_leads=np.array([.4, .6]), p_inter=np.array([[0, .5], [.5, 0]]), white_amp=1e-4, tele_p01=.01,
                    tele_p10=.02, tele_amp=.01)
    v0, dx, dy = composer.affine_physical(3, 0, -3., 1., 32, 1, -3., 1., 32)
    s = oscan.Scan(v0=v0, dx=dx, dy=dy, nx=32, ny=32, peak_width=0.1, seed=77)
    z1, n1 = oscan.simulate_scan(m, s)
    z2, n2 = oscan.simulate_scan(m, s)
    assert np.array_equal(z1, z2) and np.array_equal(n1, n2)
    s.seed = 78
    z3, _ = oscan.simulate_scan(m, s)
    assert not np.array_equal(z1, z3)


def test_pink_noise_statistics_and_spectrum():
    """The framework's own 1/f term (oracle/noise.py::pink_noise): unit-variance chains, spectrum ~ 1/f over two decades."""
    from oracle import noise
    x = noise.pink_noise(seed=12345, ny=64, nx=2048)
    assert abs(x.mean()) < 0.05 and 0.8 < x.var() < 1.2
    spec = (np.abs(np.fft.rfft(x, axis=1)) ** 2).mean(axis=0)
    f = np.fft.rfftfreq(2048)
    band = (f > 1 / 100) & (f < 1 / 6)
    slope = np.polyfit(np.log(f[band]), np.log(spec[band]), 1)[0]
    assert -1.35 < slope < -0.65, slope                      # 1/f within the band the four octaves cover
    # rows are independent chains unless carried
    y = noise.pink_noise(seed=12345, ny=4, nx=256, carry_rows=True)
    z = noise.pink_noise(seed=12345, ny=4, nx=256, carry_rows=False)
    assert np.array_equal(y[0], z[0]) and not np.allclose(y[1], z[1])
