"""The two forms of each stage of the tunnel path check each other on the GPU: lattice enumeration against block walk
(select stage, ``QDSIM_SELECT=block``), Noda iteration against Householder + multisection (eigen stage,
``QDSIM_EIGEN=householder``).  The switches are read at every launch, so one process can run all four combinations on the
same batch.  <n> is compared wherever it is well defined (spectral gap of the oracle > 1e-5 is not available at this
size; instead: the fraction of pixels that differ by more than 1e-7 is bounded as in the split-vs-mono test)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(engine, scans, pixels, n_dot):
    import torch
    from qdsim import N_F64
    z = torch.empty(pixels, dtype=torch.float32, device="cuda")
    n = torch.empty((pixels, n_dot), dtype=torch.float64, device="cuda")
    engine.scan_open(scans, z, n, N_F64, 0)
    torch.cuda.synchronize()
    assert torch.isfinite(n).all() and torch.isfinite(z).all()
    return n.cpu().numpy()


@pytest.mark.parametrize("n_dot,n_env,res", [(4, 96, 64), (5, 40, 33), (7, 24, 40), (8, 48, 64)])
def test_select_and_eigen_kernel_forms_agree(engine, n_dot, n_env, res, monkeypatch):
    from qdsim import synth
    dev = synth.sample_barrier_devices(n_env, n_dot, seed=4321 + n_dot)
    mb = synth.tunnel_batch(dev, latching=False, noise=False)
    engine.set_models(mb)
    scans = synth.env_step_scans(mb, dev, res=res, seed=17, radial=False)
    pixels = len(scans) * res * res
    out = {}
    for sel, eig in (("", ""), ("block", ""), ("", "householder"), ("block", "householder")):
        for var, val in (("QDSIM_SELECT", sel), ("QDSIM_EIGEN", eig)):
            if val:
                monkeypatch.setenv(var, val)
            else:
                monkeypatch.delenv(var, raising=False)
        out[(sel, eig)] = _run(engine, scans, pixels, n_dot)
    ref = out[("block", "householder")]
    for key, n in out.items():
        d = np.abs(n - ref).max(axis=1)
        # same 32 states and the same ground vector up to the solvers' tolerances; near-degenerate pixels (a handful in a
        # million) legitimately differ between any two solvers
        assert (d > 1e-7).mean() < 5e-5, f"{key}: {(d > 1e-7).sum()} of {pixels} pixels differ, max {d.max():.3g}"
        assert np.median(d) < 1e-9


def test_kernel_forms_agree_with_a_capacitance_model_and_points(engine, monkeypatch):
    """Voltage-dependent capacitances (the relax kernel hands s_g-scaled potentials and the cdd scale to the other stages)
    and an explicit voltage list (points mode): both forms of both stages, same numbers."""
    import torch
    from qdsim import N_F64, synth
    n_dot, n_env, res = 4, 8, 32
    dev = synth.sample_barrier_devices(n_env, n_dot, seed=77)
    mb = synth.tunnel_batch(dev, latching=False, noise=False)
    mb.params["vc_alpha"] = 0.02
    mb.params["vc_beta"] = 0.01
    engine.set_models(mb)
    scans = synth.env_step_scans(mb, dev, res=res, seed=5, radial=False)
    pixels = len(scans) * res * res
    outs = []
    for sel, eig in (("", ""), ("block", "householder")):
        for var, val in (("QDSIM_SELECT", sel), ("QDSIM_EIGEN", eig)):
            if val:
                monkeypatch.setenv(var, val)
            else:
                monkeypatch.delenv(var, raising=False)
        outs.append(_run(engine, scans, pixels, n_dot))
    d = np.abs(outs[0] - outs[1]).max(axis=1)
    assert (d > 1e-7).mean() < 1e-4 and np.median(d) < 1e-9
