"""CNN training-corpus generation on the scan kernels (rank 4 of SURVEY.md section 8f).

One call produces a whole batch of the samples that the reference makes one at a time in a thread pool
(src/qadapt/qarray_dataset/symmetric_capacitance_generator.py:108-215): random device -> target effective couplings,
symmetric, NN in [-0.7, 0.7] and NNN in [-0.3, 0.3] -> virtual gate matrix that realises them
(``_set_vgm_for_target_effective_coupling``, qarray_base_class.py:948-989) -> ground truth in those coordinates ->
gate voltages = ground truth + U[-40, 40] V, barriers = optimum + U[-r, r] -> the N-1 pair scans with radial noise,
unnormalised -> label matrix ``(N, N+1)`` in the layout the reference's dataloader reads.  All samples of a batch are one
``qd_scan_open`` launch; files are written in the reference's layout (``images/ cgd_matrices/ ground_truth/``
``batch_XXX``, :240-268).
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import maxwell, obs, synth
from ._lib import FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL
from .virtualisation import effective_coupling_vgm


def sample_targets(rng, n_samples: int, num_dots: int, coupling=(-0.7, 0.7), nnn_coupling=(-0.3, 0.3)):
    """-> (target (B, N, N) effective-coupling matrices, labels (B, N, N+1) float32).  The label holds the coupling as it
    shows in the image; the target matrix holds its negative off the diagonal (generator :137-160, 196-206)."""
    b, n = n_samples, num_dots
    nn = rng.uniform(coupling[0], coupling[1], size=(b, n - 1))
    nnn = rng.uniform(nnn_coupling[0], nnn_coupling[1], size=(b, max(n - 2, 0)))
    target = np.broadcast_to(np.eye(n), (b, n, n)).copy()
    labels = np.broadcast_to(np.eye(n, n + 1, dtype=np.float32), (b, n, n + 1)).copy()
    i = np.arange(n - 1)
    target[:, i, i + 1] = target[:, i + 1, i] = -nn
    labels[:, i, i + 1] = labels[:, i + 1, i] = nn
    j = np.arange(n - 2)
    target[:, j, j + 2] = target[:, j + 2, j] = -nnn
    labels[:, j, j + 2] = labels[:, j + 2, j] = nnn
    return target, labels


def generate_batch(eng, n_samples: int, num_dots: int, seed: int = 42, use_barriers: bool = True, res: int = 100,
                   window_delta_range=(1.5, 2.0), coupling=(-0.7, 0.7), nnn_coupling=(-0.3, 0.3),
                   voltage_offset: float = 40.0, barrier_offset_range: float = 15.0, radial: dict | None = None,
                   optimal_vg_center=(1.0, 0.53), optimal_tc: float = 1e-3, electrons: bool = True,
                   images_on_device: bool = False):
    """Generate ``n_samples`` labelled samples on ``eng``'s GPU.  Returns a dict with ``image`` (B, res, res, N-1) float32
    (the reference's per-sample layout; a CUDA tensor ``[B, N-1, res, res]`` if ``images_on_device``), ``cgd_matrix``
    (B, N, N+1) float32, ``ground_truth_voltages`` / ``gate_voltages`` (B, N) float32, ``barrier_voltages``, ``scans``."""
    import torch
    rng = np.random.default_rng(seed)
    b, n = n_samples, num_dots
    g = n + 1
    if use_barriers:
        dev = synth.sample_barrier_devices(b, n, seed=int(rng.integers(0, 2 ** 31)))
        mb = synth.tunnel_batch(dev)
    else:
        dev = synth.sample_devices(b, n, seed=int(rng.integers(0, 2 ** 31)))
        mb = synth.model_batch(dev)
    window = rng.uniform(*window_delta_range, size=b)
    target, labels = sample_targets(rng, b, n, coupling, nnn_coupling)
    cgd_gates = mb.cgd_full[:, :, :g]
    vgm = effective_coupling_vgm(mb.cdd_inv_full, cgd_gates, target, electrons=electrons)
    centre = np.concatenate([np.full(n, optimal_vg_center[0]), [optimal_vg_center[1]]])
    vg_phys = maxwell.optimal_vg(mb.cdd_inv_full, cgd_gates, centre)
    gt = np.linalg.solve(vgm, vg_phys[..., None])[..., 0][:, :n]            # origin 0 (calculate_ground_truth :1282)
    gate_v = gt + rng.uniform(-voltage_offset, voltage_offset, size=(b, n))
    barrier_v = None
    if use_barriers:
        vb_base = -np.log(optimal_tc / dev["tc_base"])[:, None] / dev["alpha"]
        vb_opt = vb_base - np.einsum("ebg,eg->eb", dev["Cbg"], vg_phys)
        barrier_v = vb_opt + rng.uniform(-barrier_offset_range, barrier_offset_range, size=vb_opt.shape)
    if radial is None:                                                       # env_config.yaml:27-35
        zero = rng.uniform(20.0, 30.0, size=b)
        radial = dict(zero_radius=zero, ramp_distance=zero + rng.uniform(5.0, 10.0, size=b),
                      full_noise_distance=rng.uniform(30.0, 40.0, size=b), max_amplitude=0.05)
    seeds = (np.uint64(seed) << np.uint64(32)) + np.arange(b * (n - 1), dtype=np.uint64)
    scans = obs.obs_scans(mb, gate_v, 0.0, vgm, np.zeros((b, g)), -window, window, res, barrier_voltages=barrier_v,
                          peak_width=dev["peak_width"], gate_ground_truth=gt, radial=radial, seeds=seeds)
    eng.set_models(mb)
    z = torch.empty(b * (n - 1) * res * res, dtype=torch.float32, device=f"cuda:{eng.device}")
    image = obs.observe(eng, scans, z, flags=FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL, normalise=False)
    if not images_on_device:
        image = image.permute(0, 2, 3, 1).contiguous().cpu().numpy()
    return {"image": image, "cgd_matrix": labels, "ground_truth_voltages": gt.astype(np.float32),
            "gate_voltages": gate_v.astype(np.float32),
            "barrier_voltages": None if barrier_v is None else barrier_v.astype(np.float32), "scans": scans,
            "virtual_gate_matrix": vgm, "target": target}


def save_batch(batch_id: int, batch: dict, output_dir: str, first_sample_id: int = 0) -> None:
    """The reference's on-disk layout (symmetric_capacitance_generator.py:240-268)."""
    for d in ("images", "cgd_matrices", "ground_truth", "metadata"):
        os.makedirs(os.path.join(output_dir, d), exist_ok=True)
    image = batch["image"]
    if not isinstance(image, np.ndarray):
        image = image.permute(0, 2, 3, 1).contiguous().cpu().numpy()
    np.save(os.path.join(output_dir, "images", f"batch_{batch_id:03d}.npy"), image)
    np.save(os.path.join(output_dir, "cgd_matrices", f"batch_{batch_id:03d}.npy"), batch["cgd_matrix"])
    gt = [{"ground_truth_voltages": batch["ground_truth_voltages"][i].tolist(),
           "gate_voltages": batch["gate_voltages"][i].tolist(), "sample_id": first_sample_id + i}
          for i in range(len(batch["cgd_matrix"]))]
    with open(os.path.join(output_dir, "ground_truth", f"batch_{batch_id:03d}.json"), "w") as f:
        json.dump(gt, f, indent=2)
