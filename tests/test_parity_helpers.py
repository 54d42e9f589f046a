"""CPU tests of the parity machinery itself (tests/util.py): the causal explanation of latched tunnel-path mismatches must
accept what it should and REJECT an unexplained difference; the C port's tie margin equals the NumPy oracle's."""
import numpy as np
import pytest

from util import explain_latched_mismatches, oracle_batch


def _case():
    rng = np.random.default_rng(0)
    ny, nx, nd = 6, 12, 4
    n_free = rng.uniform(0.0, 3.0, size=(ny, nx, nd))
    n_free[np.abs(n_free - np.floor(n_free) - 0.5) < 1e-3] += 0.01          # keep the random field off the half-integers
    gap = np.full((ny, nx), 0.1)
    return ny, nx, nd, n_free, gap


def test_identical_images_need_no_explanation():
    ny, nx, nd, n_free, gap = _case()
    d, amb, rows = explain_latched_mismatches(n_free + 3e-7, n_free, n_free, gap)
    assert (d, amb, rows) == (0, 0, ny)


def test_difference_downstream_of_a_half_integer_pixel_is_explained():
    ny, nx, nd, n_free, gap = _case()
    n_free[2, 4, 1] = 1.5 + 5e-7                     # ambiguous pixel in row 2
    gpu = n_free.copy()
    gpu[2, 6:, :] += 1.0                             # the rows' latched tail differs from pixel 6 on
    d, amb, rows = explain_latched_mismatches(gpu, n_free, n_free, gap)
    assert d == (nx - 6) and amb == 1


def test_small_gap_counts_as_ambiguous_and_flat_pass_spans_rows():
    ny, nx, nd, n_free, gap = _case()
    gap[1, 3] = 1e-7
    gpu = n_free.copy()
    gpu[4, 2, 0] += 0.5                              # differs three rows later: only a flat pass carries that far
    with pytest.raises(AssertionError):
        explain_latched_mismatches(gpu, n_free, n_free, gap)
    explain_latched_mismatches(gpu, n_free, n_free, gap, carry_rows=True)


def test_unexplained_difference_is_rejected():
    ny, nx, nd, n_free, gap = _case()
    gpu = n_free.copy()
    gpu[3, 5, 2] += 1.0
    with pytest.raises(AssertionError):
        explain_latched_mismatches(gpu, n_free, n_free, gap)
    n_free2 = n_free.copy()
    n_free2[3, 8, 0] = 2.5                           # ambiguity AFTER the first difference explains nothing
    with pytest.raises(AssertionError):
        explain_latched_mismatches(gpu, n_free2, n_free2, gap)


@pytest.mark.parametrize("alg", ["default", "thresholded", "brute_force"])
def test_cport_margin_equals_oracle_margin(alg):
    from oracle import cport
    from qdsim import synth
    dev = synth.sample_devices(2, 4, seed=5)
    mb = synth.model_batch(dev, algorithm=alg, latching=False, noise=False, threshold=0.6)
    scans = synth.env_step_scans(mb, dev, res=12, seed=6, offset_range=2.0, radial=False)
    z_ref, n_ref, margin = oracle_batch(mb, scans, 0)
    zc, nc, _, mg = cport.run_scans(mb, scans, 0, threads=2, want_margin=True)
    assert np.array_equal(nc.reshape(n_ref.shape), n_ref)
    finite = np.isfinite(margin)
    assert np.array_equal(np.isfinite(mg.reshape(margin.shape)), finite)
    np.testing.assert_allclose(mg.reshape(margin.shape)[finite], margin[finite], rtol=0, atol=1e-12)


def test_cport_parity_report_flags_a_wrong_pixel():
    from oracle import cport
    from qdsim import FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL, synth
    from util import cport_parity
    flags = FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL
    dev = synth.sample_devices(2, 4, seed=9)
    mb = synth.model_batch(dev)
    scans = synth.env_step_scans(mb, dev, res=16, seed=10)
    z, n, _ = cport.run_scans(mb, scans, flags, threads=2)
    rep = cport_parity(mb, scans, flags, z, n, threads=2)
    assert rep["ok"] and rep["n_mismatch"] == 0 and rep["z_max_abs"] == 0.0
    n2 = n.copy()
    n2[5, 1] += 1
    z2 = z.copy()
    z2[40] += 1e-3
    rep = cport_parity(mb, scans, flags, z2, n2, threads=2)
    assert not rep["ok"] and rep["n_mismatch"] >= 1 and rep["z_max_abs"] > 5e-6
