"""Batched env shell: ``QuantumDeviceEnv.reset`` / ``step`` (src/qadapt/environment/env.py:135-315) for ``n_env`` envs at
once -- first "next" row of SURVEY.md section 8f.

Everything per-env that the reference does in Python per process (device sampling, voltage ranges, rescale, reward,
ground truth, the N-1 scans, percentile normalisation) is vectorised over the env axis on the host (NumPy, O(n_env N)
numbers) or runs on the GPU (the scans and the normalisation: one ``qd_scan_open`` + one ``qd_normalise_obs`` per step
for the whole batch).  Barrier mode only, like the reference env (env.py:61-62).  Virtual-gate update methods:
``None`` (VGM stays -I for electrons, env.py:179), ``"perfect"`` (env.py:181-182) and ``"kalman"`` / ``"direct"``
(env.py:537-621): one CNN forward over the batch's ``n_env (N-1)`` device-resident scans, the batched scalar filter and
a batched pseudo-inverse (``qdsim.virtualisation``) after every observation, as the reference does per env.

Observation layout: ``image`` is a CUDA float32 tensor ``[n_env, N-1, res, res]`` (channels first; the reference's
per-env image is ``(res, res, N-1)``), voltages are NumPy ``[n_env, N]`` / ``[n_env, N-1]`` in [-1, 1].

Lifetime of an observation: the images live in TWO device buffers used alternately, so the observation returned by one
``reset`` / ``step`` stays valid across exactly one further ``step`` (``obs_t`` and ``obs_{t+1}`` never alias -- what a
rollout that stores (obs, next_obs) pairs needs; the reference returns a fresh array per step, env.py:290-291).  Keep an
observation longer than that and you must ``clone()`` it.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import maxwell, obs, synth
from ._lib import FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL


@dataclass
class EnvConfig:
    """The knobs of env_config.yaml / qarray_config.yaml this shell reads (defaults = the shipped files)."""
    max_steps: int = 50
    resolution: int = 100
    window_delta_range: tuple = (1.5, 2.0)
    constant_voltage_offset: tuple = (0.0, 0.0)
    full_plunger_range_width: tuple = (80.0, 100.0)
    full_barrier_range_width: tuple = (20.0, 30.0)
    radial_noise: dict = field(default_factory=lambda: dict(enabled=True, lower=(20.0, 30.0), ramp_range=(5.0, 10.0),
                                                            total_noise_range=(30.0, 40.0), max_amplitude=0.05))
    sparse_reward: bool = False
    plunger_radius: float = 2.0
    barrier_radius: float = 2.0
    outer_plunger_radius: float = 10.0
    outer_plunger_reward_max: float = 0.5
    gate_ramp_start: float = 40.0
    gate_quadratic_start: float = 1.0
    gate_curve_type: str = "constant"
    gate_curve_exponent: float = 2.0
    barrier_ramp_start: float = 6.0
    optimal_vg_center: tuple = (1.0, 0.53)      # dots, sensor (qarray_config.yaml:122)
    optimal_tc: float = 1e-3                    # qarray_config.yaml:125
    electrons: bool = True                      # charge_carrier_type (qarray_config.yaml:118)
    update_method: str | None = None            # None | "perfect" | "kalman" | "direct" (env_config.yaml:61)
    nearest_neighbour: bool = False             # CNN outputs [RL, LR] instead of [NN, NNN_right, NNN_left] (:62)
    variance_threshold: float = 0.05            # :65
    process_noise: float = 0.0                  # :66
    # Which simulator sits under the shell.  True (every shipped env config; env.py:61-62 refuses anything else): the
    # tunnel-coupled TunnelCoupledChargeSensed path.  False: the constant-interaction ChargeSensedDotArray path
    # (QarrayBaseClass(use_barriers=False)) with `algorithm` in default / thresholded / brute_force -- BASELINE config 3
    # ("6-dot ... brute-force ground-state search and Kalman virtualisation in the loop"); there are no barrier gates
    # then: barrier actions are ignored and barrier rewards are zero.
    use_barriers: bool = True
    algorithm: str = "default"
    max_charge_carriers: int = 4


def gate_reward(dist, cfg: EnvConfig):
    """Piecewise dense reward of env.py:420-447 / sparse reward of env.py:397-416, elementwise."""
    dist = np.asarray(dist, dtype=np.float64)
    if cfg.sparse_reward:
        out = np.zeros_like(dist)
        out[dist <= cfg.plunger_radius] = 1.0
        outer = (dist > cfg.plunger_radius) & (dist <= cfg.outer_plunger_radius)
        out[outer] = cfg.outer_plunger_reward_max * (
            1.0 - (dist[outer] - cfg.plunger_radius) / (cfg.outer_plunger_radius - cfg.plunger_radius))
        return out
    lin = 0.5 * (cfg.gate_ramp_start - dist) / (cfg.gate_ramp_start - cfg.gate_quadratic_start)
    normalized = (cfg.gate_quadratic_start - dist) / cfg.gate_quadratic_start
    if cfg.gate_curve_type == "polynomial":
        curve = np.abs(normalized) ** cfg.gate_curve_exponent
    elif cfg.gate_curve_type == "constant":
        curve = np.ones_like(dist)
    elif cfg.gate_curve_type == "exponential":
        curve = (np.exp(cfg.gate_curve_exponent * normalized) - 1) / (np.exp(cfg.gate_curve_exponent) - 1)
    elif cfg.gate_curve_type == "linear":
        curve = normalized
    else:
        raise ValueError(f"Unknown curve type: {cfg.gate_curve_type}")
    out = np.where(dist >= cfg.gate_ramp_start, 0.0, np.where(dist > cfg.gate_quadratic_start, lin, 0.5 + 0.5 * curve))
    return np.clip(out, 0.0, 1.0)


def barrier_reward(dist, cfg: EnvConfig):
    dist = np.asarray(dist, dtype=np.float64)
    if cfg.sparse_reward:
        return np.where(dist <= cfg.barrier_radius, 1.0, 0.0)
    return np.clip(np.where(dist >= cfg.barrier_ramp_start, 0.0, (cfg.barrier_ramp_start - dist) / cfg.barrier_ramp_start), 0, 1)


class BatchedDeviceEnv:
    """``n_env`` independent tuning environments stepped together on one GPU."""

    def __init__(self, n_env: int, num_dots: int, engine=None, config: EnvConfig | None = None, seed: int = 0,
                 capacitance_model=None):
        """``capacitance_model``: torch module ``images [B, 1, res, res] -> (values, log_vars) [B, K]`` (the reference's
        ``CapacitancePredictionModel`` contract); required for ``update_method`` kalman / direct."""
        self.n_env, self.num_dots = n_env, num_dots
        self.cfg = config or EnvConfig()
        self.vg_updater = None
        if self.cfg.update_method in ("kalman", "direct"):
            if capacitance_model is None:
                raise ValueError("Capacitance model weights must be provided via capacitance_model when using "
                                 f"update_method '{self.cfg.update_method}'.")
            from .virtualisation import VirtualGateUpdater
            self.vg_updater = VirtualGateUpdater(
                n_env, num_dots, capacitance_model, method=self.cfg.update_method,
                nearest_neighbour=self.cfg.nearest_neighbour, variance_threshold=self.cfg.variance_threshold,
                process_noise=self.cfg.process_noise, electrons=self.cfg.electrons)
        elif self.cfg.update_method not in (None, "perfect"):
            raise ValueError(f"Unknown update method: {self.cfg.update_method}")
        self.eng = engine
        self.rng = np.random.default_rng(seed)
        self.seed = seed
        self._episode = 0
        self.z_dev = None

    # ---- state set-up ---------------------------------------------------------------------------------------
    def _sample(self):
        E, N, cfg = self.n_env, self.num_dots, self.cfg
        rng = self.rng
        if cfg.use_barriers:
            self.dev = synth.sample_barrier_devices(E, N, seed=int(rng.integers(0, 2 ** 31)))
            self.mb = synth.tunnel_batch(self.dev)
        else:
            self.dev = synth.sample_devices(E, N, seed=int(rng.integers(0, 2 ** 31)))
            self.mb = synth.model_batch(self.dev, algorithm=cfg.algorithm, thermal=False,
                                        max_charge_carriers=cfg.max_charge_carriers)
        self.window_delta = rng.uniform(*cfg.window_delta_range, size=E)
        rn = cfg.radial_noise
        self.radial = None
        if rn and rn["enabled"]:
            zero = rng.uniform(*rn["lower"], size=E)
            self.radial = dict(zero_radius=zero, ramp_distance=zero + rng.uniform(*rn["ramp_range"], size=E),
                               full_noise_distance=rng.uniform(*rn["total_noise_range"], size=E),
                               max_amplitude=rn["max_amplitude"])
        G = N + 1
        vgm = np.broadcast_to(np.eye(G), (E, G, G)).copy()
        if cfg.update_method == "perfect":
            vgm = maxwell.optimal_vgm(self.mb.cdd_inv_full, self.mb.cgd_full[:, :, :G])
        if cfg.electrons:
            vgm = -vgm
        self.vgm = vgm
        offset = rng.uniform(*cfg.constant_voltage_offset, size=(E, N))
        self.origin = np.concatenate([offset, np.zeros((E, 1))], axis=1)

    def _ground_truth(self):
        """qarray_base_class.py:1255-1286: targets in the CURRENT virtual-gate coordinates."""
        cfg, N = self.cfg, self.num_dots
        G = N + 1
        if getattr(self, "_gt_phys_of", None) is not self.mb:     # physical optimum: a property of the device, per episode
            target = np.concatenate([np.full(N, cfg.optimal_vg_center[0]), [cfg.optimal_vg_center[1]]])
            vg_phys = maxwell.optimal_vg(self.mb.cdd_inv_full, self.mb.cgd_full[:, :, :G], target)          # (E, G)
            if cfg.use_barriers:
                tc_ratio = cfg.optimal_tc / self.dev["tc_base"]
                vb_base = -np.log(tc_ratio)[:, None] / self.dev["alpha"]
                vb_gt = vb_base - np.einsum("ebg,eg->eb", self.dev["Cbg"], vg_phys)
            else:
                vb_gt = np.zeros((self.n_env, N - 1))
            self._gt_phys = (vg_phys, vb_gt)
            self._gt_phys_of = self.mb
        vg_phys, vb = self._gt_phys
        vg_virtual = np.linalg.solve(self.vgm, (vg_phys - self.origin)[..., None])[..., 0]
        return vg_virtual[:, :-1].astype(np.float32), vb.astype(np.float32), vg_virtual[:, -1]

    def _init_voltage_ranges(self):
        """env.py:808-858."""
        E, N, cfg, rng = self.n_env, self.num_dots, self.cfg, self.rng
        pr = rng.uniform(*cfg.full_plunger_range_width, size=(E, 1))
        pc = rng.uniform(self.gate_gt - 0.5 * (pr - 2), self.gate_gt + 0.5 * (pr - 2))
        self.plunger_max, self.plunger_min = pc + 0.5 * pr, pc - 0.5 * pr
        br = rng.uniform(*cfg.full_barrier_range_width, size=(E, 1))
        bc = rng.uniform(self.barrier_gt - 0.5 * (br - 1), self.barrier_gt + 0.5 * (br - 1))
        self.barrier_max, self.barrier_min = bc + 0.5 * br, bc - 0.5 * br
        self.gate_v = rng.uniform(self.plunger_min, self.plunger_max)
        self.barrier_v = rng.uniform(self.barrier_min, self.barrier_max)

    # ---- per-step pieces ------------------------------------------------------------------------------------
    def _scans(self):
        res = self.cfg.resolution
        seeds = ((np.uint64(self.seed) << np.uint64(44)) + (np.uint64(self._episode) << np.uint64(32))
                 + (np.uint64(self.step_count) << np.uint64(20)) + np.arange(self.n_env * (self.num_dots - 1), dtype=np.uint64))
        return obs.obs_scans(self.mb, self.gate_v, self.sensor_gt, self.vgm, self.origin, -self.window_delta,
                             self.window_delta, res, barrier_voltages=self.barrier_v if self.cfg.use_barriers else None,
                             peak_width=self.dev["peak_width"],
                             gate_ground_truth=self.gate_gt if self.radial else None, radial=self.radial, seeds=seeds)

    def _reward(self):
        """env.py:350-462."""
        N = self.num_dots
        cgd_diag = np.abs(self.mb.cgd_full[:, np.arange(N), np.arange(N)])
        gate_d = np.abs(self.gate_gt - self.gate_v) * cgd_diag
        if not self.cfg.use_barriers:
            return {"gates": gate_reward(gate_d, self.cfg), "barriers": np.zeros((self.n_env, N - 1))}
        barrier_d = np.abs(self.barrier_gt - self.barrier_v) * self.dev["alpha"]
        return {"gates": gate_reward(gate_d, self.cfg), "barriers": barrier_reward(barrier_d, self.cfg)}

    def _normalised_voltages(self):
        g = ((self.gate_v.astype(np.float32) - self.plunger_min) / (self.plunger_max - self.plunger_min)) * 2 - 1
        b = ((self.barrier_v.astype(np.float32) - self.barrier_min) / (self.barrier_max - self.barrier_min)) * 2 - 1
        return g.astype(np.float32), b.astype(np.float32)

    def _observe(self):
        import torch
        res, E, N = self.cfg.resolution, self.n_env, self.num_dots
        if self.z_dev is None:
            self._z_bufs = [torch.empty(E * (N - 1) * res * res, dtype=torch.float32, device=f"cuda:{self.eng.device}")
                            for _ in range(2)]
            self._z_turn = 0
        # ping-pong: the previous observation (still held by the caller as obs_t) is not overwritten by this one
        self.z_dev = self._z_bufs[self._z_turn]
        self._z_turn ^= 1
        flags = FLAG_LATCH | FLAG_NOISE | (FLAG_RADIAL if self.radial else 0)
        image = obs.observe(self.eng, self._scans(), self.z_dev, flags=flags, normalise=True)
        if self.vg_updater is not None:                    # env.py:229 / :292: update right after the observation
            self.vgm, self.cgd_estimate = self.vg_updater.update(image, self.mb.cdd_inv_full)
        g, b = self._normalised_voltages()
        return {"image": image, "obs_gate_voltages": g, "obs_barrier_voltages": b}

    # ---- gym-like API ---------------------------------------------------------------------------------------
    def reset(self):
        self._episode += 1
        self.step_count = 0
        self._sample()
        if self.vg_updater is not None:
            self.vg_updater.reset()
        self.gate_gt, self.barrier_gt, self.sensor_gt = self._ground_truth()
        self._init_voltage_ranges()
        if self.eng is not None:
            self.eng.set_models(self.mb)
        observation = self._observe() if self.eng is not None else None
        return observation, self.info()

    def step(self, gate_actions, barrier_actions, skip_obs: bool = False):
        """Actions in [-1, 1], shapes (n_env, N) and (n_env, N-1).  Returns (obs, reward, terminated, truncated, info)."""
        self.step_count += 1
        ga = np.clip(np.asarray(gate_actions, dtype=np.float32), -1, 1)
        ba = np.clip(np.asarray(barrier_actions, dtype=np.float32), -1, 1)
        self.gate_v = (ga + 1) / 2 * (self.plunger_max - self.plunger_min) + self.plunger_min
        self.barrier_v = (ba + 1) / 2 * (self.barrier_max - self.barrier_min) + self.barrier_min
        reward = self._reward()
        truncated = np.full(self.n_env, self.step_count >= self.cfg.max_steps)
        terminated = np.zeros(self.n_env, dtype=bool)
        observation = None
        if not skip_obs and self.eng is not None:
            observation = self._observe()
            self.gate_gt, self.barrier_gt, self.sensor_gt = self._ground_truth()
        return observation, reward, terminated, truncated, self.info()

    def info(self):
        return {"gate_ground_truth": self.gate_gt, "barrier_ground_truth": self.barrier_gt,
                "sensor_ground_truth": self.sensor_gt, "current_gate_voltages": self.gate_v,
                "current_barrier_voltages": self.barrier_v, "virtual_gate_matrix": self.vgm,
                "virtual_gate_origin": self.origin}
