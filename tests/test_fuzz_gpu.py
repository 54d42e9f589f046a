"""Randomised differential test: CUDA path vs oracle over seeded random configurations -- dot count, algorithm, env
count, ragged window sizes, scattered pixel offsets, flag subsets, both entry points.  Every case is reproducible from
its integer seed (printed on failure)."""
import numpy as np
import pytest

from util import assert_z_given_n, explain_latched_mismatches, sensor_w_max, compare_charges, oracle_batch

pytestmark = pytest.mark.gpu


def _ragged(rng, scans, max_side):
    off = 0
    for rec in scans:
        nx, ny = int(rng.integers(1, max_side + 1)), int(rng.integers(1, max_side + 1))
        rec["nx"], rec["ny"], rec["pix_offset"] = nx, ny, off
        off += nx * ny + int(rng.integers(0, 5))
    return scans


@pytest.mark.parametrize("seed", range(24))
def test_path_a_random_configuration(engine, seed):
    from qdsim import (FLAG_CARRY_ROWS, FLAG_LATCH, FLAG_LATCH_EXACT, FLAG_NOISE, FLAG_RADIAL, FLAG_THERMAL,
                       FLAG_WHITE_ON_OUTPUT, N_F64, N_U8, synth)
    rng = np.random.default_rng(1000 + seed)
    n_dot = int(rng.integers(2, 9))
    alg = str(rng.choice(["default", "default", "thresholded", "brute_force"]))
    if alg == "brute_force" and n_dot > 5:
        alg = "default"
    thermal = alg != "brute_force" and rng.random() < 0.25
    n_env = int(rng.integers(1, 4))
    dev = synth.sample_devices(n_env, n_dot, seed=2000 + seed)
    mb = synth.model_batch(dev, algorithm=alg, thermal=thermal, latching=True, noise=True,
                           threshold=float(rng.uniform(0.3, 1.0)), max_charge_carriers=int(rng.integers(2, 5)))
    mb.params["tele_p01"], mb.params["tele_p10"], mb.params["tele_amp"] = 0.04, 0.09, 0.01
    engine.set_models(mb)
    scans = synth.env_step_scans(mb, dev, res=8, seed=3000 + seed, offset_range=float(rng.uniform(0.5, 4.0)))
    scans = _ragged(rng, scans[: int(rng.integers(1, min(len(scans), 4) + 1))].copy(), 45)
    scans["rad_zero_radius"], scans["rad_alpha"] = 1.0, 0.02
    if rng.random() < 0.3:
        scans["rad_mode"][-1] = 2
    flags = 0
    for f, p in ((FLAG_LATCH, 0.6), (FLAG_NOISE, 0.5), (FLAG_RADIAL, 0.5), (FLAG_CARRY_ROWS, 0.3), (FLAG_WHITE_ON_OUTPUT, 0.3)):
        if rng.random() < p:
            flags |= f
    if thermal:
        flags |= FLAG_THERMAL
        flags &= ~FLAG_LATCH_EXACT
    n_type = N_F64 if thermal else N_U8
    z, n = engine.scan_open_host(scans, n_type=n_type, flags=flags)
    noisy = bool(flags & (FLAG_NOISE | FLAG_RADIAL))
    for i, rec in enumerate(scans):
        nx, ny, o = int(rec["nx"]), int(rec["ny"]), int(rec["pix_offset"])
        z_ref, n_ref, margin = oracle_batch(mb, scans, flags, which=[i])
        zi, ni = z[o:o + nx * ny].reshape(ny, nx), n[o:o + nx * ny].reshape(ny, nx, -1)
        ctx = f"seed {seed}: N={n_dot} {alg} thermal={thermal} flags={flags:#x} scan {i} {nx}x{ny}"
        if rec["rad_mode"] == 2 and flags & FLAG_RADIAL:
            np.testing.assert_allclose(zi, z_ref[0], rtol=0, atol=5e-6, err_msg=ctx)
            continue
        if thermal:
            if flags & FLAG_LATCH:           # a rounded latch compare can flip where <n> sits on a half-integer
                bad = np.abs(ni - n_ref[0]).max(axis=-1) > 1e-8
                assert bad.mean() <= 0.02, ctx
                np.testing.assert_allclose(zi[~bad], z_ref[0][~bad], rtol=0 if noisy else 1e-6, atol=5e-6 if noisy else 1e-7, err_msg=ctx)
            else:
                np.testing.assert_allclose(ni, n_ref[0], rtol=0, atol=1e-8, err_msg=ctx)
                np.testing.assert_allclose(zi, z_ref[0], rtol=0 if noisy else 1e-6, atol=5e-6 if noisy else 1e-7, err_msg=ctx)
            continue
        if flags & FLAG_LATCH:               # a tie upstream changes what the latch sees downstream: require no ties
            if (margin[0] <= 1e-9).any():
                continue
            assert np.array_equal(ni.astype(np.int64), np.rint(n_ref[0]).astype(np.int64)), ctx
            np.testing.assert_allclose(zi, z_ref[0], rtol=0 if noisy else 1e-6, atol=5e-6 if noisy else 0, err_msg=ctx)
        else:
            compare_charges(ni, n_ref[0], margin[0], max_tie_frac=0.05)
            safe = margin[0] > 1e-9
            np.testing.assert_allclose(zi[safe], z_ref[0][safe], rtol=0 if noisy else 1e-6, atol=5e-6 if noisy else 0,
                                       err_msg=ctx)


@pytest.mark.parametrize("seed", range(10))
def test_path_b_random_configuration(engine, seed):
    from qdsim import FLAG_CARRY_ROWS, FLAG_LATCH, FLAG_NOISE, N_F64, synth
    rng = np.random.default_rng(5000 + seed)
    n_dot = int(rng.integers(4, 9))
    dev = synth.sample_barrier_devices(1, n_dot, seed=6000 + seed)
    mb = synth.tunnel_batch(dev)
    if rng.random() < 0.3:
        mb.params["vc_alpha"], mb.params["vc_beta"] = rng.uniform(0.05, 0.1), rng.uniform(0.05, 0.1)
    engine.set_models(mb)
    scans = synth.env_step_scans(mb, dev, res=8, seed=7000 + seed, offset_range=float(rng.uniform(0.5, 3.0)), radial=False)
    side = 14 if n_dot <= 6 else 7
    scans = _ragged(rng, scans[: 2].copy(), side)
    flags = 0
    for f, p in ((FLAG_LATCH, 0.5), (FLAG_NOISE, 0.5), (FLAG_CARRY_ROWS, 0.3)):
        if rng.random() < p:
            flags |= f
    z, n = engine.scan_open_host(scans, n_type=N_F64, flags=flags)
    for i, rec in enumerate(scans):
        nx, ny, o = int(rec["nx"]), int(rec["ny"]), int(rec["pix_offset"])
        z_ref, n_ref, gap = oracle_batch(mb, scans, flags, which=[i])
        zi, ni = z[o:o + nx * ny].reshape(ny, nx), n[o:o + nx * ny].reshape(ny, nx, -1)
        ctx = f"seed {seed}: N={n_dot} flags={flags:#x} scan {i} {nx}x{ny}"
        assert np.isfinite(ni).all() and np.isfinite(zi).all(), ctx
        bad = (np.abs(ni - n_ref[0]).max(axis=-1) > 2e-6) | (gap[0] <= 1e-5)
        if flags & FLAG_LATCH:
            # causal: a differing pixel must lie downstream (same latching sequence) of a pixel whose free <n> is within
            # the solver tolerance of a half-integer, or whose spectral gap leaves the ground vector ill-conditioned
            _, n_free, _ = oracle_batch(mb, scans, flags & ~FLAG_LATCH, which=[i])
            explain_latched_mismatches(ni, n_ref[0], n_free[0], gap[0], carry_rows=bool(flags & FLAG_CARRY_ROWS),
                                       n_atol=2e-6, half_tol=4e-6)
        else:
            assert (bad & (gap[0] > 1e-5)).sum() == 0, ctx
        assert_z_given_n(zi, z_ref[0], ni, n_ref[0], ~bad, sensor_w_max(mb, 0), float(rec["peak_width"]),
                         noise_atol=5e-6 if flags & FLAG_NOISE else 0.0, what=ctx)
