"""The plain-C restatement (oracle/cport) agrees with the NumPy oracle on seeded inputs.  CPU only."""
import numpy as np
import pytest

from util import oracle_batch


@pytest.mark.parametrize("n_dot,alg,flag_names", [
    (2, "default", ()), (4, "default", ("LATCH", "NOISE", "RADIAL")), (4, "default", ("LATCH", "NOISE", "CARRY_ROWS")),
    (6, "default", ("LATCH",)), (5, "thresholded", ()), (3, "brute_force", ()), (4, "default", ("THERMAL",)),
    (3, "default", ("NOISE", "WHITE_ON_OUTPUT")),
])
def test_cport_matches_numpy_oracle(n_dot, alg, flag_names):
    import qdsim
    from oracle import cport
    from qdsim import synth
    flags = 0
    for f in flag_names:
        flags |= getattr(qdsim, "FLAG_" + f)
    dev = synth.sample_devices(2, n_dot, seed=40 + n_dot)
    mb = synth.model_batch(dev, algorithm=alg, thermal="THERMAL" in flag_names, threshold=0.6)
    mb.params["tele_p01"], mb.params["tele_p10"], mb.params["tele_amp"] = 0.04, 0.07, 0.01
    scans = synth.env_step_scans(mb, dev, res=24, seed=5, offset_range=3.0)
    scans["rad_zero_radius"], scans["rad_alpha"] = 1.0, 0.02
    z, n, _ = cport.run_scans(mb, scans, flags, threads=2)
    z_ref, n_ref, _ = oracle_batch(mb, scans, flags)
    np.testing.assert_allclose(n.reshape(n_ref.shape), n_ref, rtol=0, atol=1e-10)
    np.testing.assert_allclose(z.reshape(z_ref.shape), z_ref, rtol=3e-7, atol=1e-7)
