"""Virtual-gate update in the loop, for a whole env batch (rank 2 of SURVEY.md section 8f; BASELINE config 3's
"Kalman virtualisation in the loop").

What the reference does per env and per step (src/qadapt/environment/env.py:537-621): push the N-1 normalised scans
through a CNN that returns ``(values, log_vars)`` per scan, feed them -- negated, "qarray sign convention" -- into a
scalar Kalman filter per capacitive coupling with variance gating (src/qadapt/capacitance_model/KalmanUpdater.py:92-213)
or into the direct updater (DirectUpdater.py), read the estimated Cgd back (``get_full_matrix``, :222-227) and turn it
into a new virtual gate matrix (src/qadapt/environment/qarray_base_class.py:904-942).

Here the same arithmetic runs once for ``n_env`` envs:

* ``BatchedCapacitanceUpdater`` -- the Kalman / direct filter with a leading env axis.  The reference's update ORDER is
  kept (scan 0..N-2; within a scan NN, NNN-right, NNN-left or RL, LR) because successive scans hit the same
  next-nearest-neighbour element; each of those <= 3(N-1) sequential steps is vectorised over envs.  Bit-identical to
  the reference classes (``tests/test_virtualisation.py`` runs the real ones side by side).
* ``virtual_gate_matrices`` -- ``-pinv(cdd_inv_full @ cgd_estimate_full)`` for the batch (one batched SVD).
* ``CapacitanceCNN`` -- the reference's ``CapacitancePredictionModel`` architecture (MobileNetV3-small trunk, 1-channel
  stem, value + log-variance heads; CapacitancePrediction.py:114-202) with the same module names, so the reference's
  checkpoints load with ``load_state_dict``.  Stock torch layers; runs on the images where they already are (HBM).
* ``VirtualGateUpdater`` -- glue: images ``[E, N-1, res, res]`` on the device -> CNN (one forward for all scans of all
  envs) -> filter -> new ``vgm (E, G, G)``.
"""
from __future__ import annotations

import numpy as np


class BatchedCapacitanceUpdater:
    """``n_env`` independent ``KalmanCapacitanceUpdater`` / ``DirectCapacitanceUpdater`` states."""

    def __init__(self, n_env: int, n_dots: int, method: str = "kalman", prior_mean: float = 0.0,
                 prior_variance: float = 0.5, variance_threshold: float = 0.05, process_noise: float = 0.0,
                 include_nnn: bool = True, mean_bounds=(-1.0, 1.0), log_var_bounds=(-6.0, 2.0),
                 prior_mean_nnn: float | None = None):
        if method not in ("kalman", "direct"):
            raise ValueError(f"Unknown update method: {method}")
        self.n_env, self.n_dots, self.method = n_env, n_dots, method
        self.variance_threshold, self.process_noise = variance_threshold, process_noise
        self.prior_mean = prior_mean
        self.prior_mean_nnn = prior_mean_nnn if prior_mean_nnn is not None else prior_mean
        self.prior_variance = prior_variance
        self.include_nnn = include_nnn
        self.mean_bounds, self.log_var_bounds = mean_bounds, log_var_bounds
        self.means = np.zeros((n_env, n_dots, n_dots))
        self.variances = np.zeros((n_env, n_dots, n_dots))
        self.total_accepted = np.zeros(n_env, dtype=np.int64)
        self.total_rejected = np.zeros(n_env, dtype=np.int64)
        self.reset()

    def reset(self, env_mask=None):
        """Back to the prior (all envs, or those selected by a boolean mask -- envs of a batch reset independently)."""
        sel = slice(None) if env_mask is None else np.asarray(env_mask, dtype=bool)
        i = np.arange(self.n_dots - 1)
        self.means[sel] = 0.0
        self.variances[sel] = 0.0
        m, v = self.means[sel], self.variances[sel]
        m[:, i, i + 1] = m[:, i + 1, i] = self.prior_mean
        v[:, i, i + 1] = v[:, i + 1, i] = self.prior_variance
        if self.include_nnn:
            j = np.arange(self.n_dots - 2)
            m[:, j, j + 2] = m[:, j + 2, j] = self.prior_mean_nnn
            v[:, j, j + 2] = v[:, j + 2, j] = self.prior_variance
        self.means[sel], self.variances[sel] = m, v
        self.total_accepted[sel] = 0
        self.total_rejected[sel] = 0

    def _variance(self, log_var):
        return np.exp(np.clip(log_var, self.log_var_bounds[0], self.log_var_bounds[1]))

    def update(self, i: int, j: int, delta, measurement_variance):
        """One coupling (i, j) of every env: ``delta`` / ``measurement_variance`` (E,).  Returns the accepted mask."""
        row, col = min(i, j), max(i, j)
        delta = np.asarray(delta, dtype=np.float64)
        r = np.asarray(measurement_variance, dtype=np.float64)
        ok = ~(r > self.variance_threshold)
        if self.method == "kalman":
            p = self.variances[:, row, col] + self.process_noise
            x = self.means[:, row, col]
            k = p / (p + r)
            new_mean = x + k * delta
            new_var = (1 - k) * p
        else:
            new_mean, new_var = delta, r
        new_mean = np.clip(new_mean, self.mean_bounds[0], self.mean_bounds[1])
        self.means[:, row, col] = self.means[:, col, row] = np.where(ok, new_mean, self.means[:, row, col])
        self.variances[:, row, col] = self.variances[:, col, row] = np.where(ok, new_var, self.variances[:, row, col])
        self.total_accepted += ok
        self.total_rejected += ~ok
        return ok

    def update_from_scans(self, deltas, log_vars):
        """``deltas``, ``log_vars``: (E, N-1, K) -- K = 3 outputs [NN, NNN_right, NNN_left] or K = 2 [RL, LR] per scan
        (``update_from_scan`` of the reference for left_dot = 0..N-2, in that order)."""
        deltas = np.asarray(deltas, dtype=np.float64)
        var = self._variance(np.asarray(log_vars, dtype=np.float64))
        n = self.n_dots
        k_out = deltas.shape[-1]
        assert deltas.shape == (self.n_env, n - 1, k_out) and var.shape == deltas.shape
        for i in range(n - 1):
            if self.include_nnn and k_out == 3:
                self.update(i, i + 1, deltas[:, i, 0], var[:, i, 0])
                if i + 2 < n:
                    self.update(i, i + 2, deltas[:, i, 1], var[:, i, 1])
                if i - 1 >= 0:
                    self.update(i + 1, i - 1, deltas[:, i, 2], var[:, i, 2])
            elif k_out == 2:
                self.update(i + 1, i, deltas[:, i, 0], var[:, i, 0])
                self.update(i, i + 1, deltas[:, i, 1], var[:, i, 1])
            else:
                raise ValueError(f"Expected 2 or 3 outputs, got {k_out}")

    def get_capacitance_stats(self, i: int, j: int):
        return self.means[:, i, j], self.variances[:, i, j]

    def get_full_matrix(self):
        """(E, N, N) estimated Cgd with the diagonal set to 1."""
        cgd = self.means.copy()
        d = np.arange(self.n_dots)
        cgd[:, d, d] = 1.0
        return cgd


def virtual_gate_matrices(cdd_inv_full, cgd_estimate, electrons: bool = True, cbd=None):
    """Barrier-mode ``QarrayBaseClass._update_virtual_gate_matrix`` (qarray_base_class.py:904-942) for a batch.

    ``cdd_inv_full`` (E, D, D) with D = N + 1; ``cgd_estimate`` (E, N, N) positive plunger-to-dot couplings.  The
    estimate is embedded as ``[[cgd_est, 0], [0, 1]]`` (sensor gate column 0, sensor coupling 1), negated (Maxwell sign),
    and ``vgm = -pinv(cdd_inv_full @ cgd_gates)``, sign-flipped for electrons.  ``cbd`` (barrier columns, only used when
    virtualising barriers) does not enter the gate-only pseudo-inverse, exactly as in the reference.
    """
    cdd_inv_full = np.asarray(cdd_inv_full, dtype=np.float64)
    est = np.asarray(cgd_estimate, dtype=np.float64)
    e, n = est.shape[0], est.shape[-1]
    full = np.zeros((e, n + 1, n + 1))
    full[:, :n, :n] = est
    full[:, n, n] = 1.0
    vgm = -np.linalg.pinv(cdd_inv_full @ (-full))
    return -vgm if electrons else vgm


def effective_coupling_vgm(cdd_inv_full, cgd_gates, target, electrons: bool = True):
    """``_set_vgm_for_target_effective_coupling`` (qarray_base_class.py:948-989), batched: VGM with
    ``cdd_inv cgd VGM = T`` for a target effective-coupling matrix ``target`` (E, N, N) (used by the dataset generator)."""
    target = np.asarray(target, dtype=np.float64)
    e, n = target.shape[0], target.shape[-1]
    t_full = np.broadcast_to(np.eye(n + 1), (e, n + 1, n + 1)).copy()
    t_full[:, :n, :n] = target
    vgm = -np.linalg.pinv(np.asarray(cdd_inv_full) @ np.asarray(cgd_gates)) @ t_full
    return -vgm if electrons else vgm


def make_capacitance_cnn(output_size: int = 3, mobilenet: str = "small"):
    """The reference's ``CapacitancePredictionModel`` (CapacitancePrediction.py:114-202): torchvision MobileNetV3 trunk
    with a 1-channel stem and identity classifier, ``value_head`` and ``confidence_head`` MLPs (feature -> 256 -> 128 ->
    outputs, ReLU + Dropout 0.2).  Module names match, so reference checkpoints load; weights here are random (there is
    no network for the ImageNet initialisation and no checkpoint in the reference tree)."""
    import torch.nn as nn
    from torchvision import models

    class CapacitanceCNN(nn.Module):
        def __init__(self):
            super().__init__()
            if mobilenet == "small":
                self.backbone, feat = models.mobilenet_v3_small(weights=None), 576
            elif mobilenet == "large":
                self.backbone, feat = models.mobilenet_v3_large(weights=None), 960
            else:
                raise ValueError(mobilenet)
            stem = self.backbone.features[0][0]
            self.backbone.features[0][0] = nn.Conv2d(1, stem.out_channels, stem.kernel_size, stem.stride, stem.padding,
                                                     bias=stem.bias is not None)
            self.backbone.classifier = nn.Identity()
            self.output_size = output_size

            def head():
                return nn.Sequential(nn.Linear(feat, 256), nn.ReLU(), nn.Dropout(0.2), nn.Linear(256, 128), nn.ReLU(),
                                     nn.Dropout(0.2), nn.Linear(128, output_size))
            self.value_head, self.confidence_head = head(), head()

        def forward(self, x):
            f = self.backbone(x)
            return self.value_head(f), self.confidence_head(f)

    return CapacitanceCNN()


class VirtualGateUpdater:
    """CNN + filter + VGM for a batch of envs (env.py:537-621 with ``update_method`` kalman | direct)."""

    def __init__(self, n_env: int, n_dots: int, ml_model, method: str = "kalman", nearest_neighbour: bool = False,
                 variance_threshold: float = 0.05, process_noise: float = 0.0, electrons: bool = True,
                 chunk: int = 8192, autocast_dtype=None):
        self.n_env, self.n_dots = n_env, n_dots
        self.ml_model = ml_model
        self.nearest_neighbour = nearest_neighbour
        self.electrons = electrons
        self.chunk = chunk
        self.autocast_dtype = autocast_dtype
        # priors of env.py:779-787
        self.predictor = BatchedCapacitanceUpdater(
            n_env, n_dots, method=method, prior_mean=0.3, prior_variance=0.5, variance_threshold=variance_threshold,
            process_noise=process_noise, include_nnn=not nearest_neighbour, prior_mean_nnn=0.15)

    def reset(self, env_mask=None):
        self.predictor.reset(env_mask)

    def predict(self, image):
        """``image``: torch tensor [E, N-1, H, W] (any device) -> (values, log_vars) NumPy (E, N-1, K)."""
        import torch
        e, c, h, w = image.shape
        x = image.reshape(e * c, 1, h, w).float()
        dev = next(self.ml_model.parameters()).device
        vals, lvs = [], []
        with torch.no_grad():
            for s in range(0, x.shape[0], self.chunk):
                xb = x[s:s + self.chunk].to(dev)
                if self.autocast_dtype is not None:
                    with torch.autocast(device_type=dev.type, dtype=self.autocast_dtype):
                        v, lv = self.ml_model(xb)
                else:
                    v, lv = self.ml_model(xb)
                vals.append(v.float())
                lvs.append(lv.float())
        values = torch.cat(vals).reshape(e, c, -1).cpu().numpy()
        log_vars = torch.cat(lvs).reshape(e, c, -1).cpu().numpy()
        return values, log_vars

    def update(self, image, cdd_inv_full):
        """One in-loop update: returns the new ``vgm (E, G, G)`` and the Cgd estimate ``(E, N, N)``."""
        values, log_vars = self.predict(image)
        self.predictor.update_from_scans(-values.astype(np.float64), log_vars.astype(np.float64))   # env.py:596-618
        est = self.predictor.get_full_matrix()
        return virtual_gate_matrices(cdd_inv_full, est, self.electrons), est
