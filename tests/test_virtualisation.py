"""Virtual-gate update in the loop against the reference's own classes.

``tests/golden/ref_virtualisation.npz`` was produced by the reference's ``KalmanCapacitanceUpdater`` /
``DirectCapacitanceUpdater`` (src/qadapt/capacitance_model/*.py, imported as they are) driven by the caller's loop of
env.py:596-618, and by ``QarrayBaseClass._update_virtual_gate_matrix`` / ``_set_vgm_for_target_effective_coupling``
(qarray_base_class.py:904-989) -- see tests/golden/make_reference_golden.py::run_reference_virtualisation.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(HERE, "golden", "ref_virtualisation.npz")))


@pytest.mark.parametrize("method", ["kalman", "direct"])
@pytest.mark.parametrize("k_out", [3, 2])
def test_batched_updater_is_bit_identical_to_reference(gold, method, k_out):
    from qdsim.virtualisation import BatchedCapacitanceUpdater
    tag = f"{method}_k{k_out}"
    values, log_vars = gold[tag + "_values"], gold[tag + "_log_vars"]
    n_step, n_env, n_pairs, _ = values.shape
    upd = BatchedCapacitanceUpdater(n_env, n_pairs + 1, method=method, prior_mean=0.3, prior_variance=0.5,
                                    variance_threshold=0.05, process_noise=np.where(np.arange(n_env) % 2, 0.01, 0.0),
                                    include_nnn=(k_out == 3), prior_mean_nnn=0.15)
    for t in range(n_step):
        upd.update_from_scans(-values[t].astype(np.float64), log_vars[t].astype(np.float64))
        assert np.array_equal(upd.means, gold[tag + "_means"][t])
        assert np.array_equal(upd.variances, gold[tag + "_variances"][t])
        assert np.array_equal(upd.get_full_matrix(), gold[tag + "_full"][t])
    assert (upd.total_accepted > 0).all() and (upd.total_rejected > 0).all()       # the gate is exercised both ways
    upd.reset(np.arange(n_env) == 1)
    fresh = BatchedCapacitanceUpdater(1, n_pairs + 1, method=method, prior_mean=0.3, prior_mean_nnn=0.15,
                                      include_nnn=(k_out == 3))
    assert np.array_equal(upd.means[1], fresh.means[0]) and np.array_equal(upd.means[0], gold[tag + "_means"][-1][0])


def test_vgm_update_matches_reference(gold):
    from qdsim.virtualisation import effective_coupling_vgm, virtual_gate_matrices
    g = int(gold["n_dot"]) + 1
    for electrons, key in ((True, "vgm_electrons"), (False, "vgm_holes")):
        vgm = virtual_gate_matrices(gold["cdd_inv_full"], gold["cgd_estimate"], electrons=electrons)
        np.testing.assert_allclose(vgm, gold[key], rtol=1e-11, atol=1e-13)
    vgm = effective_coupling_vgm(gold["cdd_inv_full"], gold["cgd_full"][:, :, :g], gold["target"], electrons=False)
    np.testing.assert_allclose(vgm, gold["vgm_target_holes"], rtol=1e-11, atol=1e-13)


def test_capacitance_cnn_contract():
    """Same module names / shapes as the reference's CapacitancePredictionModel, so its checkpoints load."""
    import torch
    from qdsim.virtualisation import make_capacitance_cnn
    m = make_capacitance_cnn(output_size=3).eval()
    keys = set(m.state_dict())
    assert "backbone.features.0.0.weight" in keys and m.state_dict()["backbone.features.0.0.weight"].shape[1] == 1
    assert {"value_head.0.weight", "value_head.6.weight", "confidence_head.6.bias"} <= keys
    assert m.state_dict()["value_head.0.weight"].shape == (256, 576)
    with torch.no_grad():
        v, lv = m(torch.zeros(4, 1, 64, 64))
    assert v.shape == (4, 3) and lv.shape == (4, 3)


def test_updater_glue_cpu():
    """VirtualGateUpdater end to end on CPU tensors with a stub CNN: negation, order and VGM shape."""
    import torch
    from qdsim import synth
    from qdsim.virtualisation import BatchedCapacitanceUpdater, VirtualGateUpdater, virtual_gate_matrices

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        def forward(self, x):
            m = x.mean(dim=(1, 2, 3))
            return torch.stack([m, 2 * m, 3 * m], dim=1), torch.full((x.shape[0], 3), -5.0)

    e, n, res = 3, 4, 8
    mb = synth.tunnel_batch(synth.sample_barrier_devices(e, n, seed=1))
    image = torch.rand(e, n - 1, res, res)
    u = VirtualGateUpdater(e, n, Stub())
    vgm, est = u.update(image, mb.cdd_inv_full)
    ref = BatchedCapacitanceUpdater(e, n, prior_mean=0.3, prior_mean_nnn=0.15)
    m = image.mean(dim=(2, 3)).numpy().astype(np.float64)
    ref.update_from_scans(-np.stack([m, 2 * m, 3 * m], axis=-1), np.full((e, n - 1, 3), -5.0))
    assert np.array_equal(est, ref.get_full_matrix())
    np.testing.assert_allclose(vgm, virtual_gate_matrices(mb.cdd_inv_full, est), rtol=0, atol=0)
    assert vgm.shape == (e, n + 1, n + 1)


def test_env_requires_model_for_kalman():
    from qdsim.vector_env import BatchedDeviceEnv, EnvConfig
    with pytest.raises(ValueError):
        BatchedDeviceEnv(2, 4, config=EnvConfig(update_method="kalman"))
    with pytest.raises(ValueError):
        BatchedDeviceEnv(2, 4, config=EnvConfig(update_method="bayesian"))


@pytest.mark.gpu
def test_batched_env_with_kalman_in_the_loop(engine):
    """BASELINE config 3's shape at small size: env rollout with CNN + Kalman virtualisation in the loop, on cuda:0."""
    import torch
    from qdsim.vector_env import BatchedDeviceEnv, EnvConfig
    from qdsim.virtualisation import make_capacitance_cnn
    torch.manual_seed(0)
    cnn = make_capacitance_cnn(3).cuda().eval()
    e, n = 8, 4
    env = BatchedDeviceEnv(e, n, engine=engine, config=EnvConfig(resolution=32, update_method="kalman"), seed=3,
                           capacitance_model=cnn)
    obs, info = env.reset()
    assert obs["image"].shape == (e, n - 1, 32, 32) and obs["image"].is_cuda
    vgm0 = env.vgm.copy()
    assert not np.allclose(vgm0, -np.eye(n + 1))                 # the reset's first update already moved it (env.py:229)
    rng = np.random.default_rng(0)
    for _ in range(3):
        obs, reward, term, trunc, info = env.step(rng.uniform(-1, 1, (e, n)), rng.uniform(-1, 1, (e, n - 1)))
    assert np.isfinite(env.vgm).all() and env.vgm.shape == (e, n + 1, n + 1)
    acc, rej = env.vg_updater.predictor.total_accepted, env.vg_updater.predictor.total_rejected
    assert ((acc + rej) == 4 * (3 * (n - 1) - 2)).all()           # 4 updates x (NN + NNN-right + NNN-left per scan, edges cut)
    # the scans really use the updated matrix: descriptors of the next step carry its columns
    s = env._scans()
    np.testing.assert_allclose(s["dx"][0, :n + 1], env.vgm[0][:, 0] * (2 * env.window_delta[0] / 31), rtol=1e-12)
