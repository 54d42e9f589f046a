"""Path B on the GPU (QD_ALG_TUNNEL: what QADAPT's env.step runs in barrier mode) against the literal NumPy restatement
of src/qarray_latched/DotArrays/ground_state.py (oracle/path_b.py).

<n> is an eigenvector expectation: it is compared within 1e-6 absolute wherever the pixel's spectral gap exceeds 1e-5
(below that the ground vector itself is ill-conditioned: LAPACK and any other solver legitimately differ)."""
import numpy as np
import pytest

from util import assert_z_given_n, explain_latched_mismatches, oracle_batch, sensor_w_max

pytestmark = pytest.mark.gpu

N_ATOL = 1e-6
GAP_MIN = 1e-5


def _assert_z(z, z_ref, n, n_ref, ok, mb, scans, noisy=False):
    """z at 1e-6 relative given <n> (tests/util.py::assert_z_given_n), per scan (peak width and env vary)."""
    z_ref = np.asarray(z_ref)
    z, n, ok = np.asarray(z).reshape(z_ref.shape), np.asarray(n).reshape(np.asarray(n_ref).shape), np.asarray(ok).reshape(z_ref.shape)
    for i, rec in enumerate(scans):
        assert_z_given_n(z[i], z_ref[i], n[i], n_ref[i], ok[i], sensor_w_max(mb, int(rec["env_id"])),
                         float(rec["peak_width"]), noise_atol=5e-6 if noisy else 0.0, what=f"scan {i}:")


def _setup(engine, n_dot, n_env, res, seed, **kw):
    from qdsim import synth
    dev = synth.sample_barrier_devices(n_env, n_dot, seed=seed)
    mb = synth.tunnel_batch(dev, **kw)
    engine.set_models(mb)
    scans = synth.env_step_scans(mb, dev, res=res, seed=seed + 1, offset_range=2.5)
    return dev, mb, scans


@pytest.mark.parametrize("n_dot,res,n_env", [(4, 32, 2), (5, 24, 1), (6, 16, 1), (8, 8, 1)])
def test_tunnel_ground_state_matches_oracle(engine, n_dot, res, n_env):
    from qdsim import N_F64
    dev, mb, scans = _setup(engine, n_dot, n_env, res, seed=20 + n_dot, latching=False, noise=False)
    scans = scans[: min(len(scans), 3)].copy()
    scans["pix_offset"] = np.arange(len(scans)) * res * res
    z, n = engine.scan_open_host(scans, n_type=N_F64, flags=0)
    z_ref, n_ref, gap = oracle_batch(mb, scans, 0)
    z, n = z.reshape(z_ref.shape), n.reshape(n_ref.shape)
    ok = gap > GAP_MIN
    assert ok.mean() > 0.9
    assert np.isfinite(n).all() and np.isfinite(z).all()
    np.testing.assert_allclose(n[ok], n_ref[ok], rtol=0, atol=N_ATOL)
    _assert_z(z, z_ref, n, n_ref, ok, mb, scans)
    assert (np.abs(n_ref - np.rint(n_ref)) > 0.05).any(), "tunnel coupling should smear some transitions"


def test_tunnel_with_latching_and_noise(engine):
    from qdsim import FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL, N_F64
    flags = FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL
    dev, mb, scans = _setup(engine, 4, 2, 24, seed=31)
    scans["rad_zero_radius"], scans["rad_alpha"] = 1.0, 0.02
    z, n = engine.scan_open_host(scans, n_type=N_F64, flags=flags)
    z_ref, n_ref, gap = oracle_batch(mb, scans, flags)
    _, n_free, _ = oracle_batch(mb, scans, flags & ~FLAG_LATCH)
    z, n = z.reshape(z_ref.shape), n.reshape(n_ref.shape)
    # every differing pixel must be EXPLAINED: downstream (same row) of a pixel whose free <n> sits on a half-integer
    # within the eigen-solver tolerance, or whose spectral gap makes the ground vector ill-conditioned
    n_dif = n_amb = n_rows = 0
    for i in range(len(scans)):
        d, a, r = explain_latched_mismatches(n[i], n_ref[i], n_free[i], gap[i])
        n_dif, n_amb, n_rows = n_dif + d, n_amb + a, n_rows + r
    assert n_amb <= 0.05 * n_rows, f"{n_amb} of {n_rows} rows ambiguous: the test has no teeth"
    same = np.abs(n - n_ref).max(axis=-1) <= N_ATOL
    _assert_z(z, z_ref, n, n_ref, same, mb, scans, noisy=True)


def test_tunnel_points_mode_flat_pass(engine):
    """The facade's barrier-mode call: charge_sensor_open(vg_flat (P, G), vb (P, B)) -- one row of P points."""
    from oracle import composer
    from qdsim import FLAG_LATCH, N_F64
    dev, mb, scans = _setup(engine, 4, 1, 16, seed=41, noise=False)
    rec = scans[0]
    nv = mb.n_volt
    v = composer.affine_grid(rec["v0"][:nv], rec["dx"][:nv], rec["dy"][:nv], 16, 16).reshape(1, 256, nv)
    z, n = engine.points_open_host(rec, v, n_type=N_F64, flags=FLAG_LATCH)
    flat = scans[:1].copy()
    from qdsim import FLAG_CARRY_ROWS
    z_ref, n_ref, gap = oracle_batch(mb, flat, FLAG_LATCH | FLAG_CARRY_ROWS)
    _, n_free, _ = oracle_batch(mb, flat, 0)
    explain_latched_mismatches(n.reshape(16, 16, -1), n_ref[0], n_free[0], gap[0], carry_rows=True)
    same = np.abs(n.reshape(16, 16, -1) - n_ref[0]).max(axis=-1) <= N_ATOL
    assert same.mean() > 0.5
    _assert_z(z.reshape(1, 16, 16), z_ref, n.reshape(1, 16, 16, -1), n_ref, same[None], mb, flat)


def test_tunnel_constant_tc_without_barriers(engine):
    """No barrier voltages: constant nearest-neighbour coupling model.tc (ground_state.py:92-101); tc -> 0 recovers the
    integer argmin over the kept states (SURVEY.md section 8c test 5)."""
    from qdsim import N_F64, synth
    from qdsim.engine import ModelBatch, PARAMS_DTYPE
    from qdsim import maxwell
    dev = synth.sample_devices(1, 4, seed=51)
    cdd_nm, cgd_nm = maxwell.embed_sensor(dev["Cdd"], dev["Cgd"], dev["Cds"], dev["Cgs"])
    _, cdi_f, cgd_f = maxwell.maxwell(cdd_nm, cgd_nm)
    params = np.zeros(1, dtype=PARAMS_DTYPE)
    params["tc_base"] = 1e-9
    mb = ModelBatch(algorithm="tunnel", n_gate=5, cdd_inv_gs=np.ascontiguousarray(cdi_f[:, :4, :4]), cdd_gs=None,
                    cdd_inv_full=cdi_f, cgd_full=cgd_f, params=params)
    engine.set_models(mb)
    scans = synth.env_step_scans(mb, dev, res=24, seed=52, offset_range=2.0, radial=False)[:1]
    z, n = engine.scan_open_host(scans, n_type=N_F64, flags=0)
    z_ref, n_ref, gap = oracle_batch(mb, scans, 0)
    ok = gap.reshape(-1) > 1e-4
    assert np.abs(n[ok] - np.rint(n[ok])).max() < 1e-6, "t -> 0: occupations are integers away from degeneracies"
    np.testing.assert_allclose(n[ok], n_ref.reshape(-1, 4)[ok], rtol=0, atol=N_ATOL)


@pytest.mark.parametrize("n_dot", [6, 8])
def test_tunnel_one_wide_sector(engine, n_dot):
    """A stiff common mode (adding a carrier costs far more than moving one): all 32 kept states share one total charge,
    i.e. ONE sector of 32 lanes -- the widest case of the segmented solver (segments longer than 16 lanes)."""
    from qdsim import N_F64
    from qdsim.engine import ModelBatch, PARAMS_DTYPE, new_scans
    n, d = n_dot, n_dot + 1
    rng = np.random.default_rng(70 + n_dot)
    a = np.diag(rng.uniform(0.04, 0.07, size=d))
    cinv_full = (a + 1.5 * np.ones((d, d)))[None]
    cgd_full = -np.concatenate([np.eye(d) + 0.02 * rng.uniform(size=(d, d)), 0.01 * rng.uniform(size=(d, n - 1))], axis=1)[None]
    params = np.zeros(1, dtype=PARAMS_DTYPE)
    params["tc_base"] = 0.03
    params["alpha"][0, :n - 1] = 1.0
    mb = ModelBatch(algorithm="tunnel", n_gate=d, cdd_inv_gs=np.ascontiguousarray(cinv_full[:, :n, :n]), cdd_gs=None,
                    cdd_inv_full=cinv_full, cgd_full=cgd_full, params=params, cbg=np.zeros((1, n - 1, d)))
    engine.set_models(mb)
    res = 8
    s = new_scans(1)
    s["v0"][0, :d] = -rng.uniform(1.2, 2.8, size=d)
    s["dx"][0, 0], s["dy"][0, 1] = -0.6 / res, -0.6 / res
    s["nx"], s["ny"], s["peak_width"] = res, res, 0.2
    z, nn = engine.scan_open_host(s, n_type=N_F64, flags=0)
    z_ref, n_ref, gap = oracle_batch(mb, s, 0)
    # the construction really yields a single sector
    from oracle import composer, path_b
    from util import oracle_model
    m = oracle_model(mb, 0, 0)
    v = composer.affine_grid(s["v0"][0, :mb.n_volt], s["dx"][0, :mb.n_volt], s["dy"][0, :mb.n_volt], res, res).reshape(-1, mb.n_volt)
    g = v @ np.asarray(m.cgd).T
    st = path_b.select_charge_states(g, path_b.continuous_ground_state(g, m.cdd_inv), m.cdd_inv, 32, 1000)
    widths = np.array([np.unique(t, return_counts=True)[1].max() for t in st.sum(axis=-1)])
    assert (widths > 16).mean() > 0.5, widths
    ok = gap.reshape(-1) > GAP_MIN
    assert ok.mean() > 0.8
    np.testing.assert_allclose(nn.reshape(-1, n)[ok], n_ref.reshape(-1, n)[ok], rtol=0, atol=N_ATOL)
    _assert_z(z.reshape(z_ref.shape), z_ref, nn.reshape(n_ref.shape), n_ref, ok.reshape(z_ref.shape), mb, s)


@pytest.mark.parametrize("n_dot,res,n_env", [(4, 24, 3), (6, 16, 2), (8, 12, 2), (2, 16, 1), (3, 16, 1)])
def test_split_pipeline_equals_monolithic_kernel(engine, n_dot, res, n_env, monkeypatch):
    """The three-kernel pipeline (relax / select / eigen) against the single-kernel form it was cut from: the same 32 states
    and the same solver, so <n> agrees far inside the solver tolerance (the relaxation's summation order differs)."""
    from qdsim import FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL, N_F64
    dev, mb, scans = _setup(engine, n_dot, n_env, res, seed=90 + n_dot)
    scans["rad_mode"][::3] = 2
    flags = FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL
    out = {}
    for mono in ("1", "0"):
        monkeypatch.setenv("QDSIM_TUNNEL_MONO", mono)
        out[mono] = engine.scan_open_host(scans, n_type=N_F64, flags=0), engine.scan_open_host(scans, n_type=N_F64, flags=flags)
    _, _, gap = oracle_batch(mb, scans, 0)
    ok = gap.reshape(-1) > GAP_MIN
    (z1, n1), (z1f, n1f) = out["1"]
    (z0, n0), (z0f, n0f) = out["0"]
    assert np.isfinite(n0).all() and np.isfinite(z0).all()
    np.testing.assert_allclose(n0[ok], n1[ok], rtol=0, atol=1e-9)
    np.testing.assert_allclose(z0[ok], z1[ok], rtol=1e-7, atol=1e-9)
    same = np.abs(n0f - n1f).max(axis=-1) <= 1e-9
    assert same.mean() > 0.97
    np.testing.assert_allclose(z0f[same], z1f[same], rtol=1e-6, atol=1e-8)


def test_split_pipeline_chunking(engine, monkeypatch):
    """More scans than one scratch chunk holds (QDSIM_TUNNEL_CHUNK_PIX): chunk boundaries must not show."""
    from qdsim import FLAG_LATCH, FLAG_NOISE, N_F64
    dev, mb, scans = _setup(engine, 4, 6, 32, seed=97)                      # 18 scans of 1024 pixels
    flags = FLAG_LATCH | FLAG_NOISE
    z_all, n_all = engine.scan_open_host(scans, n_type=N_F64, flags=flags)
    for pix in ("2048", "5000", "1"):                                       # 2, 4 and 1 scans per chunk
        monkeypatch.setenv("QDSIM_TUNNEL_CHUNK_PIX", pix)
        z, n = engine.scan_open_host(scans, n_type=N_F64, flags=flags)
        np.testing.assert_array_equal(n, n_all)
        np.testing.assert_array_equal(z, z_all)


@pytest.mark.parametrize("n_dot,n_env", [(4, 400), (8, 96)])
def test_large_batch_exercises_the_warm_started_walk(engine, n_dot, n_env, monkeypatch):
    """Batches large enough that a work item is a run of 64 consecutive pixels: the select kernel's warm start, live-block
    pick and second-level bound only run there (small batches give every pixel its own item).  No NaN / inf anywhere, and
    the split pipeline equals the single-kernel form (cold and warm walks select the same 32 states)."""
    import torch
    from qdsim import FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL, N_F64, synth
    dev = synth.sample_barrier_devices(n_env, n_dot, seed=1234)
    mb = synth.tunnel_batch(dev)
    engine.set_models(mb)
    scans = synth.env_step_scans(mb, dev, res=64, seed=99)
    pixels = len(scans) * 4096
    assert pixels >= 148 * 16 * 4 * 64                       # items of 64 pixels (qd_api.cu: want_items = SMs * 16 * 4)
    flags = FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL
    out = {}
    for mono in ("0", "1"):
        monkeypatch.setenv("QDSIM_TUNNEL_MONO", mono)
        z = torch.empty(pixels, dtype=torch.float32, device="cuda")
        n = torch.empty((pixels, n_dot), dtype=torch.float64, device="cuda")
        engine.scan_open(scans, z, n, N_F64, 0)              # unlatched <n>: pixel-wise comparable
        torch.cuda.synchronize()
        assert torch.isfinite(n).all() and torch.isfinite(z).all(), f"mono={mono}: {(~torch.isfinite(n)).sum().item()} bad values"
        out[mono] = (z, n)
        zf = torch.empty(pixels, dtype=torch.float32, device="cuda")
        engine.scan_open(scans, zf, None, 0, flags)
        torch.cuda.synchronize()
        assert torch.isfinite(zf).all()
    dn = (out["0"][1] - out["1"][1]).abs().max(dim=1).values
    # the two forms use the same solver; they differ by the relaxation's summation order (floor flips on exact integers) and
    # by near-degenerate pixels -- a handful in a million
    assert (dn > 1e-7).float().mean().item() < 2e-5, f"{(dn > 1e-7).sum().item()} of {pixels} pixels differ"
