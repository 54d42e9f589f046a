"""The reference-facing classes (``qarray.ChargeSensedDotArray`` ...) on the GPU, written the way the reference calls
them (src/qadapt/environment/qarray_base_class.py:128-139, 744-756), checked against the CPU oracle."""
import numpy as np
import pytest

from oracle import capacitance as cap
from oracle import composer, path_a, sensor

pytestmark = pytest.mark.gpu

CDD = [[0, .12], [.12, 0]]
CGD = [[1.0, .35, 0], [.3, .97, 0]]
CDS = [[.04, .045]]
CGS = [[6e-5, 3e-5, .98]]


def _model(**kw):
    import qarray
    return qarray.ChargeSensedDotArray(Cdd=CDD, Cgd=CGD, Cds=CDS, Cgs=CGS, coulomb_peak_width=0.15, T=0.0,
                                       algorithm="default", implementation="jax", max_charge_carriers=4, **kw)


def test_baseline_config1_2dot_64x64_do2d_open_bit_exact():
    """BASELINE.json configs[0]: 2-dot ChargeSensedDotArray, single 64x64 do2d_open scan, noise-free."""
    m = _model()
    z, n = m.do2d_open(1, -3.3, 0.7, 64, 2, -3.1, 0.9, 64)
    assert z.shape == (64, 64, 1) and n.shape == (64, 64, 2) and z.dtype == np.float64
    vg = composer.do2d(3, 1, -3.3, 0.7, 64, 2, -3.1, 0.9, 64)
    cdd, cdi, cgd = cap.convert_to_maxwell(CDD, CGD)
    _, cdi_f, cgd_f = cap.with_sensor(CDD, CGD, CDS, CGS)
    n_ref, margin = path_a.ground_state_open(vg.reshape(-1, 3), cgd, cdi, cdd, return_margin=True)
    safe = (margin > 1e-9).reshape(64, 64)
    assert safe.mean() > 0.999
    assert np.array_equal(n[safe], n_ref.reshape(64, 64, 2)[safe])
    z_ref = sensor.charge_sensor_signal(n_ref, vg.reshape(-1, 3), cdi_f, cgd_f, 0.15).reshape(64, 64, 1)
    np.testing.assert_allclose(z[safe], z_ref[safe], rtol=1e-6)
    assert n.max() >= 3 and n.min() == 0


def test_ground_state_and_sensor_on_explicit_voltage_arrays():
    m = _model()
    rng = np.random.default_rng(0)
    vg = rng.uniform(-3, 0.5, (7, 50, 3))
    n = m.ground_state_open(vg)
    assert n.shape == (7, 50, 2)
    cdd, cdi, cgd = cap.convert_to_maxwell(CDD, CGD)
    n_ref, margin = path_a.ground_state_open(vg.reshape(-1, 3), cgd, cdi, cdd, return_margin=True)
    ok = margin.reshape(7, 50) > 1e-9
    assert np.array_equal(n[ok], n_ref.reshape(7, 50, 2)[ok])
    z, n2 = m.charge_sensor_open(vg)
    assert z.shape == (7, 50, 1) and np.array_equal(n2, n)
    # virtual gates: the default virtual gate matrix decouples the dots
    zv, nv = m.do2d_open("vP1", 0.2, 2.8, 32, "vP2", 0.2, 2.8, 32)
    assert np.array_equal(nv[:, :, 0], np.broadcast_to(nv[0, :, 0], (32, 32)))
    # mutable peak width takes effect without re-upload
    m.coulomb_peak_width = 0.4
    z_wide, _ = m.charge_sensor_open(vg)
    assert not np.allclose(z, z_wide)


def test_error_behaviour_matches_the_reference():
    import qarray
    m = _model()
    with pytest.raises(ValueError):
        m.ground_state_open(np.zeros((4, 5)))
    with pytest.raises(AssertionError):
        qarray.ChargeSensedDotArray(Cdd=CDD, Cgd=CGD, Cds=CDS, Cgs=CGS, algorithm="thresholded", implementation="jax")
    with pytest.raises(AssertionError):
        qarray.ChargeSensedDotArray(Cdd=CDD, Cgd=CGD, Cds=CDS, Cgs=CGS, algorithm="nope")
    with pytest.raises(AssertionError):
        m.optimal_Vg([1, 1])
    with pytest.raises(ValueError):
        qarray.ChargeSensedDotArray(Cdd=[[0, -.1], [-.1, 0]], Cgd=CGD, Cds=CDS, Cgs=CGS)


def test_noise_and_latching_models_are_honoured_and_seedable():
    import qarray
    noise = qarray.WhiteNoise(amplitude=5e-3) + qarray.TelegraphNoise(p01=0.02, p10=0.05, amplitude=0.02)
    latch = qarray.LatchingModel(n_dots=2, p_leads=[0.3, 0.3], p_inter=[[0, 0.3], [0.3, 0]])
    quiet = _model()
    loud = _model(noise_model=noise, latching_model=latch)
    np.random.seed(5)
    z1, n1 = loud.do2d_open(1, -3.3, 0.7, 48, 2, -3.1, 0.9, 48)
    np.random.seed(5)
    z2, n2 = loud.do2d_open(1, -3.3, 0.7, 48, 2, -3.1, 0.9, 48)
    z3, n3 = loud.do2d_open(1, -3.3, 0.7, 48, 2, -3.1, 0.9, 48)
    z0, n0 = quiet.do2d_open(1, -3.3, 0.7, 48, 2, -3.1, 0.9, 48)
    assert np.array_equal(z1, z2) and np.array_equal(n1, n2)
    assert not np.array_equal(z1, z3)
    assert (n1 != n0).any(), "latching with p = 0.3 must delay some transitions"
    assert 1e-5 < np.abs(z1 - z0)[n1.sum(-1) == n0.sum(-1)].std() < 1.0
    thermal = _model()
    thermal.T = 100.0
    nt = thermal.ground_state_open(np.array([[-0.52, -0.5, 0.0]]))
    assert nt.shape == (1, 2)


def test_tunnel_coupled_class_as_the_facade_calls_it():
    """qarray_base_class.py:817-838 constructor kwargs and :143-163 call pattern: composer.do2d('vP1', ..., 'vP2', ...,
    gate_voltages, True) -> flatten -> charge_sensor_open(vg_flat, vb)."""
    import qarray
    from oracle import capacitance as ocap
    from oracle import path_b, sensor
    from oracle import scan as oscan
    from qarray_latched.DotArrays.barrier_voltage_model import BarrierVoltageModel
    from qarray_latched.DotArrays.TunnelCoupledChargeSensed import TunnelCoupledChargeSensed
    from qdsim import synth
    dev = synth.sample_barrier_devices(1, 4, seed=61)
    raw = {k: dev[k][0] for k in ("Cdd", "Cgd", "Cds", "Cgs", "Cbd", "Cbg", "Cbs")}
    bm = BarrierVoltageModel(n_barrier=3, n_dot=4, tc_base=float(dev["tc_base"][0]), alpha=list(dev["alpha"][0]))
    m = TunnelCoupledChargeSensed(**raw, Cbb=np.eye(3), barrier_model=bm, coulomb_peak_width=0.2, T=100.0,
                                  max_charge_carriers=4, tc=0.15, noise_model=None, latching_model=None,
                                  voltage_capacitance_model=None, use_sparse=False, num_charge_states=32,
                                  charge_state_batch_size=1000, charge_carrier="electrons")
    res = 24
    m.gate_voltage_composer.virtual_gate_matrix = -np.eye(5)
    gv = np.array([0.6, 0.9, 0.4, 0.7, 0.5])
    vg = m.gate_voltage_composer.do2d("vP1", gv[0] - 1.5, gv[0] + 1.5, res, "vP2", gv[1] - 1.5, gv[1] + 1.5, res, gv, True)
    vg_flat = vg.reshape(-1, vg.shape[-1])
    vb = np.full((vg_flat.shape[0], 3), np.array([1.5, 2.0, 0.5]))
    z, n = m.charge_sensor_open(vg_flat, vb)
    assert z.shape == (res * res, 1) and n.shape == (res * res, 4)
    # oracle on the same explicit voltages
    _, cdi_f, cgd_f = ocap.with_barriers_and_sensor(raw["Cdd"], raw["Cgd"], raw["Cds"], raw["Cgs"], raw["Cbd"], raw["Cbs"])
    om = oscan.Model(cdd_inv=cdi_f[:4, :4], cdd=None, cgd=cgd_f[:4], cdd_inv_full=cdi_f, cgd_full=cgd_f,
                     algorithm="tunnel", n_gate=5, cbg=raw["Cbg"], tc_base=bm.tc_base, alpha=np.asarray(bm.alpha))
    v_ext = np.concatenate([vg_flat, vb], axis=1)
    n_ref, gap = path_b.ground_state_open(om, v_ext, return_gap=True)
    ok = gap > 1e-5
    assert ok.mean() > 0.9
    np.testing.assert_allclose(n[ok], n_ref[ok], rtol=0, atol=1e-6)
    z_ref = sensor.charge_sensor_signal(n_ref, v_ext, cdi_f, cgd_f, 0.2)
    from util import assert_z_given_n
    assert_z_given_n(z.reshape(-1), z_ref.reshape(-1), n, n_ref, ok, float(np.abs(cdi_f[4, :4]).max()), 0.2)
    with pytest.raises(ValueError):
        m.charge_sensor_open(vg_flat)            # barrier voltages are required for a model built with barriers
