"""Multi-agent face of the batched env (rank 3 of SURVEY.md section 8f): ``MultiAgentEnvWrapper``
(src/qadapt/environment/multi_agent_wrapper.py:100-584) for ``n_env`` envs at once, without the RLlib base class.

Same agents (``plunger_0..N-1``, ``barrier_0..N-2``), same action combination (``_combine_agent_actions``, :386-425),
same reward distribution (``_distribute_rewards``, :427-457), same per-agent infos (``ground_truth`` /
``current_voltage``, :546-570), same ``__all__`` termination keys -- every value carries a leading env axis, and the
per-agent images are zero-copy views of the batch's CUDA image (``qdsim.agents``).  ``episode_stats`` packs the
quantities the reference logs per episode (return, length, final distances: metrics_logger.py:71-99) for the one
collective of the design, ``parallel.gather_episode_stats``.
"""
from __future__ import annotations

import numpy as np

from . import agents
from .vector_env import BatchedDeviceEnv


class BatchedMultiAgentEnv:
    def __init__(self, base_env: BatchedDeviceEnv, return_voltage: bool = True):
        self.base_env = base_env
        self.return_voltage = return_voltage
        self.num_gates = base_env.num_dots
        self.num_barriers = base_env.num_dots - 1
        self.num_image_channels = base_env.num_dots - 1
        self.gate_agent_ids = [f"plunger_{i}" for i in range(self.num_gates)]
        self.barrier_agent_ids = [f"barrier_{i}" for i in range(self.num_barriers)]
        self.all_agent_ids = self.gate_agent_ids + self.barrier_agent_ids
        # channel assignment of multi_agent_wrapper.py:147-178
        self.agent_channel_map = {}
        for i, aid in enumerate(self.gate_agent_ids):
            last = self.num_gates - 2
            self.agent_channel_map[aid] = [0, 0] if i == 0 else [last, last] if i == self.num_gates - 1 else [i - 1, i]
        for j, aid in enumerate(self.barrier_agent_ids):
            self.agent_channel_map[aid] = [j]
        self._returns = None
        self._lengths = None

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _combine_agent_actions(self, agent_actions: dict):
        e = self.base_env.n_env
        gate = np.zeros((e, self.num_gates), dtype=np.float32)
        barrier = np.zeros((e, self.num_barriers), dtype=np.float32)
        for ids, out in ((self.gate_agent_ids, gate), (self.barrier_agent_ids, barrier)):
            for aid in ids:
                if aid in agent_actions:
                    a = np.asarray(agent_actions[aid], dtype=np.float32).reshape(e, -1)
                    out[:, int(aid.split("_")[1])] = a[:, 0]
        return gate, barrier

    def _distribute_rewards(self, rewards: dict) -> dict:
        if "gates" not in rewards:
            raise ValueError("Missing gate rewards in global_rewards")
        if "barriers" not in rewards:
            raise ValueError("Missing barrier rewards in global_rewards")
        out = {aid: rewards["gates"][:, i] for i, aid in enumerate(self.gate_agent_ids)}
        out.update({aid: rewards["barriers"][:, j] for j, aid in enumerate(self.barrier_agent_ids)})
        return out

    def _observations(self, obs):
        if obs is None:
            return None
        return agents.agent_observations(obs, self.num_gates, return_voltage=self.return_voltage)

    def _infos(self, info: dict) -> dict:
        out = {}
        for i, aid in enumerate(self.gate_agent_ids):
            out[aid] = {"ground_truth": info["gate_ground_truth"][:, i], "current_voltage": info["current_gate_voltages"][:, i]}
        for j, aid in enumerate(self.barrier_agent_ids):
            out[aid] = {"ground_truth": info["barrier_ground_truth"][:, j],
                        "current_voltage": info["current_barrier_voltages"][:, j]}
        return out

    # ---- MultiAgentEnv-shaped API ------------------------------------------------------------------------------
    def reset(self):
        obs, info = self.base_env.reset()
        e = self.base_env.n_env
        self._returns = np.zeros(e)
        self._lengths = np.zeros(e, dtype=np.int64)
        self._last_info = info
        return self._observations(obs), {aid: info for aid in self.all_agent_ids}

    def step(self, agent_actions: dict, skip_obs: bool = False):
        assert len(agent_actions) == len(self.all_agent_ids), "Agent actions must match the number of agents"
        assert all(aid in self.agent_channel_map for aid in agent_actions), "Unknown agent IDs in actions"
        gate, barrier = self._combine_agent_actions(agent_actions)
        obs, rewards, terminated, truncated, info = self.base_env.step(gate, barrier, skip_obs=skip_obs)
        agent_rewards = self._distribute_rewards(rewards)
        self._returns += rewards["gates"].sum(axis=1) + rewards["barriers"].sum(axis=1)
        self._lengths += 1
        self._last_info = info
        agent_terminated = {aid: terminated for aid in self.all_agent_ids}
        agent_terminated["__all__"] = bool(np.all(terminated))
        agent_truncated = {aid: truncated for aid in self.all_agent_ids}
        agent_truncated["__all__"] = bool(np.all(truncated))
        return self._observations(obs), agent_rewards, agent_terminated, agent_truncated, self._infos(info)

    def episode_stats(self) -> np.ndarray:
        """(n_env, 4) float32: episode return, length, mean |plunger distance|, mean |barrier distance| -- the shard's
        contribution to ``parallel.gather_episode_stats``."""
        info = self._last_info
        gd = np.abs(info["current_gate_voltages"] - info["gate_ground_truth"]).mean(axis=1)
        bd = np.abs(info["current_barrier_voltages"] - info["barrier_ground_truth"]).mean(axis=1)
        return np.stack([self._returns, self._lengths.astype(np.float64), gd, bd], axis=1).astype(np.float32)
