"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle).

CPU: the oracle and its C port reproduce the committed fixtures.  GPU: the CUDA path reproduces them through the C ABI.
"""
import numpy as np
import pytest

from util import (GOLDEN_CASES, GOLDEN_TUNNEL_CASES, assert_z_given_n, compare_charges, explain_latched_mismatches,
                  load_golden, oracle_batch, sensor_w_max)


@pytest.mark.parametrize("name", GOLDEN_CASES + GOLDEN_TUNNEL_CASES)
def test_oracle_reproduces_golden(name):
    mb, scans, flags, z, n, margin = load_golden(name)
    z2, n2, m2 = oracle_batch(mb, scans, flags)
    if name in GOLDEN_TUNNEL_CASES:                       # eigenvector expectation: LAPACK build to LAPACK build
        ok = margin > 1e-6
        np.testing.assert_allclose(n2[ok], n[ok], rtol=0, atol=1e-9)
        np.testing.assert_allclose(z2[ok], z[ok], rtol=1e-8, atol=1e-10)
        return
    assert np.array_equal(n2, n)
    np.testing.assert_allclose(z2, z, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cport_reproduces_golden(name):
    from oracle import cport
    from qdsim import FLAG_THERMAL
    mb, scans, flags, z, n, margin = load_golden(name)
    zc, nc, _ = cport.run_scans(mb, scans, flags, threads=2)
    if flags & FLAG_THERMAL:
        np.testing.assert_allclose(nc.reshape(n.shape), n, rtol=0, atol=1e-10)
    else:
        assert np.array_equal(nc.reshape(n.shape), n)
    np.testing.assert_allclose(zc.reshape(z.shape), z, rtol=3e-7, atol=1e-7)      # C port stores fp32 images


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_gpu_reproduces_golden(engine, name):
    from qdsim import FLAG_NOISE, FLAG_RADIAL, FLAG_THERMAL, N_F64, N_U8
    mb, scans, flags, z, n, margin = load_golden(name)
    engine.set_models(mb)
    thermal = bool(flags & FLAG_THERMAL)
    zg, ng = engine.scan_open_host(scans, n_type=N_F64 if thermal else N_U8, flags=flags)
    zg, ng = zg.reshape(z.shape), ng.reshape(n.shape)
    if thermal:
        np.testing.assert_allclose(ng, n, rtol=0, atol=1e-9)
    else:
        compare_charges(ng, n, margin)
    noisy = bool(flags & (FLAG_NOISE | FLAG_RADIAL))
    np.testing.assert_allclose(zg, z, rtol=0 if noisy else 1e-6, atol=5e-6 if noisy else 1e-7 if thermal else 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_TUNNEL_CASES)
def test_gpu_reproduces_tunnel_golden(engine, name):
    from qdsim import FLAG_NOISE, N_F64
    mb, scans, flags, z, n, gap = load_golden(name)
    engine.set_models(mb)
    zg, ng = engine.scan_open_host(scans, n_type=N_F64, flags=flags)
    from qdsim import FLAG_CARRY_ROWS, FLAG_LATCH
    zg, ng = zg.reshape(z.shape), ng.reshape(n.shape)
    same = (np.abs(ng - n).max(axis=-1) <= 1e-6) & (gap > 1e-5)
    if flags & FLAG_LATCH:
        # every differing pixel must be downstream of a half-integer <n> / small-gap pixel of its own row
        _, n_free, _ = oracle_batch(mb, scans, flags & ~FLAG_LATCH)
        amb = rows = 0
        for i in range(len(scans)):
            _, a, r = explain_latched_mismatches(ng[i], n[i], n_free[i], gap[i], carry_rows=bool(flags & FLAG_CARRY_ROWS))
            amb, rows = amb + a, rows + r
        assert amb <= 0.05 * rows, f"{amb} of {rows} latching sequences ambiguous"
    else:
        assert same[gap > 1e-5].all(), f"{(~same & (gap > 1e-5)).sum()} pixels differ"
    for i, rec in enumerate(scans):
        assert_z_given_n(zg[i], z[i], ng[i], n[i], same[i], sensor_w_max(mb, int(rec["env_id"])), float(rec["peak_width"]),
                         noise_atol=5e-6 if flags & FLAG_NOISE else 0.0, what=f"{name} scan {i}:")
