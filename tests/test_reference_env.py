"""The REAL reference ``QuantumDeviceEnv`` (src/qadapt/environment/env.py) running reset() / step() UNCHANGED on top of
our drop-in ``qarray`` / ``qarray_latched`` packages (north_star: "src/qadapt's env ... run unchanged on top").
Only in this container (needs /root/reference); the engine is the CPU oracle here, see tests/ref_harness.py."""
import os

import numpy as np
import pytest
import yaml

import ref_harness

pytestmark = pytest.mark.skipif(not os.path.isdir(ref_harness.REF), reason="reference tree not present")


def _config(tmp_path, num_dots=4, res=10, update_method=None):
    cfg = yaml.safe_load(open(os.path.join(ref_harness.REF, "qadapt/environment/env_config.yaml")))
    cfg["simulator"]["num_dots"] = num_dots
    cfg["simulator"]["resolution"] = res
    cfg["simulator"]["max_steps"] = 3
    cfg["capacitance_model"]["update_method"] = update_method
    path = tmp_path / "env_config.yaml"
    path.write_text(yaml.safe_dump(cfg))
    return str(path)


# update_method "fake" is bit-rotted in the reference itself for barrier models: fake_capacitance_model returns the
# (N, N+1) model.cgd, which QarrayBaseClass._update_virtual_gate_matrix then fails to hstack (qarray_base_class.py:922);
# "kalman" / "direct" need the CNN checkpoint.
@pytest.mark.parametrize("update_method", [None, "perfect"])
def test_reference_env_reset_and_step_run_unchanged(monkeypatch, tmp_path, update_method):
    base, envmod, eng = ref_harness.install(monkeypatch)
    import qarray_latched
    np.random.seed(3)
    env = envmod.QuantumDeviceEnv(config_path=_config(tmp_path, update_method=update_method))
    assert isinstance(env.array.model, qarray_latched.TunnelCoupledChargeSensed)
    obs, info = env.reset()
    assert obs["image"].shape == (10, 10, 3) and obs["image"].dtype == np.float32
    assert 0.0 <= obs["image"].min() and obs["image"].max() <= 1.0
    launches = eng.launch_count
    for _ in range(3):
        action = {"action_gate_voltages": np.random.uniform(-1, 1, 4), "action_barrier_voltages": np.random.uniform(-1, 1, 3)}
        obs, reward, terminated, truncated, info = env.step(action)
        assert obs["image"].shape == (10, 10, 3) and np.isfinite(obs["image"]).all()
        assert reward["gates"].shape == (4,) and reward["barriers"].shape == (3,)
        assert obs["obs_gate_voltages"].shape == (4,) and np.abs(obs["obs_gate_voltages"]).max() <= 1 + 1e-6
    assert truncated and not terminated
    assert eng.launch_count == launches + 9            # three steps x (N-1) scans through charge_sensor_open(vg_flat, vb)
    state = info["current_device_state"]
    assert state["virtual_gate_matrix"].shape == (5, 5) and len(state["gate_ground_truth"]) == 4
