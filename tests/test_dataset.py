"""Batched dataset generation (reference: qarray_dataset/symmetric_capacitance_generator.py)."""
import json
import os

import numpy as np
import pytest


def test_targets_and_labels_follow_the_generator_layout():
    from qdsim.dataset import sample_targets
    rng = np.random.default_rng(0)
    target, labels = sample_targets(rng, 7, 5)
    assert target.shape == (7, 5, 5) and labels.shape == (7, 5, 6) and labels.dtype == np.float32
    assert np.array_equal(target, target.transpose(0, 2, 1))
    i = np.arange(4)
    assert (np.abs(labels[:, i, i + 1]) <= 0.7).all() and (np.abs(labels[:, np.arange(3), np.arange(3) + 2]) <= 0.3).all()
    np.testing.assert_allclose(target[:, i, i + 1], -labels[:, i, i + 1].astype(np.float64), rtol=1e-6)   # image shows -target
    np.testing.assert_allclose(labels[:, i + 1, i], labels[:, i, i + 1])
    assert (labels[:, :, 5] == 0).all() and (labels[:, np.arange(5), np.arange(5)] == 1).all()
    assert (target[:, 0, 3] == 0).all() and (target[:, 0, 4] == 0).all()


def test_vgm_realises_the_target_coupling():
    """cdd_inv cgd VGM = +-T: the defining property of _set_vgm_for_target_effective_coupling (values pinned against the
    reference in tests/test_virtualisation.py)."""
    from qdsim import synth
    from qdsim.dataset import sample_targets
    from qdsim.virtualisation import effective_coupling_vgm
    b, n = 4, 6
    mb = synth.tunnel_batch(synth.sample_barrier_devices(b, n, seed=2))
    target, _ = sample_targets(np.random.default_rng(1), b, n)
    a = mb.cdd_inv_full @ mb.cgd_full[:, :, :n + 1]
    vgm = effective_coupling_vgm(mb.cdd_inv_full, mb.cgd_full[:, :, :n + 1], target, electrons=True)
    t_full = np.broadcast_to(np.eye(n + 1), (b, n + 1, n + 1)).copy()
    t_full[:, :n, :n] = target
    np.testing.assert_allclose(a @ vgm, t_full, rtol=0, atol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("use_barriers", [False, True])
def test_generate_and_save_batch(engine, tmp_path, use_barriers):
    from qdsim.dataset import generate_batch, save_batch
    b, n, res = 6, 4, 32
    batch = generate_batch(engine, b, n, seed=5, use_barriers=use_barriers, res=res, voltage_offset=3.0)
    assert batch["image"].shape == (b, res, res, n - 1) and batch["image"].dtype == np.float32
    assert np.isfinite(batch["image"]).all() and batch["image"].std() > 1e-3
    again = generate_batch(engine, b, n, seed=5, use_barriers=use_barriers, res=res, voltage_offset=3.0)
    assert np.array_equal(batch["image"], again["image"])              # counter-based RNG: same seed, same corpus
    other = generate_batch(engine, b, n, seed=6, use_barriers=use_barriers, res=res, voltage_offset=3.0)
    assert not np.array_equal(batch["image"], other["image"])
    save_batch(3, batch, str(tmp_path), first_sample_id=18)
    img = np.load(os.path.join(tmp_path, "images", "batch_003.npy"))
    cgd = np.load(os.path.join(tmp_path, "cgd_matrices", "batch_003.npy"))
    gt = json.load(open(os.path.join(tmp_path, "ground_truth", "batch_003.json")))
    assert img.shape == (b, res, res, n - 1) and cgd.shape == (b, n, n + 1)
    assert len(gt) == b and gt[0]["sample_id"] == 18 and len(gt[0]["gate_voltages"]) == n


@pytest.mark.gpu
def test_far_windows_are_replaced_by_noise(engine):
    """+-40 V offsets put most windows beyond full_noise_distance: those scans are pure N(0,1) (qarray_base_class.py:463-468)."""
    from qdsim.dataset import generate_batch
    batch = generate_batch(engine, 16, 4, seed=9, res=32, voltage_offset=80.0)
    s = batch["scans"]
    far = s["rad_mode"] == 2
    assert far.mean() > 0.3
    img = batch["image"].transpose(0, 3, 1, 2).reshape(-1, 32 * 32)
    assert abs(img[far].std() - 1.0) < 0.05 and abs(img[far].mean()) < 0.05
