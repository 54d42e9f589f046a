"""RLlib-facing adaptor: the ``MultiAgentEnv`` contract of the reference's
``MultiAgentEnvWrapper`` (src/qadapt/environment/multi_agent_wrapper.py:459-584) on top of the batched CUDA env, so that
``training/train.py:532-539`` (``register_env`` + ``env_creator``) can consume it.

    reset(*, seed=None, options=None) -> (obs_dict, info_dict)
    step(action_dict)                 -> (obs_dict, reward_dict, terminated_dict, truncated_dict, info_dict)

* agent ids ``plunger_0..N-1`` / ``barrier_0..N-2`` (:113-116); ``__all__`` keys in the terminated / truncated dicts (:527-531);
* per-agent observation ``{'image': (H, W, 2 | 1) float32, 'voltage': (1,) float32}`` with the channel assignment and
  transposes of ``_extract_agent_observation`` (:311-383) -- optionally ``global_image`` / ``global_voltages`` for
  centralised critics (``return_global_state``);
* per-agent infos ``{'ground_truth', 'current_voltage'}`` (:546-570);
* ``observation_space`` / ``action_space`` dicts keyed by agent id, ``_agent_ids`` / ``agents`` / ``possible_agents`` (:300-309).

``n_env == 1`` is the drop-in: values have exactly the reference's shapes (NumPy).  ``n_env > 1`` is the vector form for a
custom env runner: the SAME dict structure with a leading env axis on every leaf (``image`` ``(E, H, W, C)``), still one
batched launch per step; ``sub_env(e)`` slices out env ``e`` in the reference's shapes.  The class derives from
``ray.rllib.env.multi_agent_env.MultiAgentEnv`` when ray is importable and is a plain duck-typed class otherwise (this image
has no ray); spaces are ``gymnasium.spaces`` when gymnasium is importable, light stand-ins with ``shape / low / high / dtype``
otherwise.
"""
from __future__ import annotations

import numpy as np

from .multi_agent import BatchedMultiAgentEnv
from .vector_env import BatchedDeviceEnv, EnvConfig

try:                                                   # pragma: no cover - ray is absent from the build image
    from ray.rllib.env.multi_agent_env import MultiAgentEnv as _Base
except Exception:                                      # noqa: BLE001
    class _Base:                                       # duck-typed stand-in: RLlib only needs the methods below
        pass

try:                                                   # pragma: no cover - gymnasium is absent from the build image
    from gymnasium import spaces as _spaces

    def _box(low, high, shape):
        return _spaces.Box(low=low, high=high, shape=shape, dtype=np.float32)

    def _dict(**kw):
        return _spaces.Dict(**kw)
except Exception:                                      # noqa: BLE001
    class _BoxStub:
        def __init__(self, low, high, shape):
            self.low = np.full(shape, low, dtype=np.float32)
            self.high = np.full(shape, high, dtype=np.float32)
            self.shape, self.dtype = tuple(shape), np.float32

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool((x >= self.low).all() and (x <= self.high).all())

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(np.float32)

    class _DictStub(dict):
        @property
        def spaces(self):
            return self

    def _box(low, high, shape):
        return _BoxStub(low, high, shape)

    def _dict(**kw):
        return _DictStub(**kw)


class VectorMultiAgentEnv(_Base):
    def __init__(self, n_env: int = 1, num_dots: int = 4, engine=None, config: EnvConfig | None = None, seed: int = 0,
                 capacitance_model=None, return_voltage: bool = True, return_global_state: bool = False,
                 to_numpy: bool = True, base_env: BatchedDeviceEnv | None = None):
        try:
            super().__init__()
        except TypeError:
            pass
        self.base_env = base_env if base_env is not None else BatchedDeviceEnv(
            n_env, num_dots, engine=engine, config=config, seed=seed, capacitance_model=capacitance_model)
        self._ma = BatchedMultiAgentEnv(self.base_env, return_voltage=return_voltage)
        self.n_env = self.base_env.n_env
        self.return_voltage = return_voltage
        self.return_global_state = return_global_state
        self.to_numpy = to_numpy
        self.num_gates = self._ma.num_gates
        self.num_barriers = self._ma.num_barriers
        self.num_image_channels = self._ma.num_image_channels
        self.use_barriers = True
        self.gate_agent_ids = list(self._ma.gate_agent_ids)
        self.barrier_agent_ids = list(self._ma.barrier_agent_ids)
        self.all_agent_ids = list(self._ma.all_agent_ids)
        self.agent_channel_map = dict(self._ma.agent_channel_map)
        self._agent_ids = set(self.all_agent_ids)
        self.agents = self._agent_ids.copy()
        self.possible_agents = self._agent_ids.copy()
        self._create_agent_spaces()

    # ---- spaces (multi_agent_wrapper.py:180-309) ---------------------------------------------------------------
    def _create_agent_spaces(self):
        res = self.base_env.cfg.resolution
        lead = () if self.n_env == 1 else (self.n_env,)
        n_glob = self.num_gates + self.num_barriers
        obs, act = {}, {}
        for aid in self.all_agent_ids:
            ch = len(self.agent_channel_map[aid])
            img = _box(0.0, 1.0, lead + (res, res, ch))
            if self.return_voltage:
                d = {"image": img, "voltage": _box(-1.0, 1.0, lead + (1,))}
                if self.return_global_state:
                    d["global_image"] = _box(0.0, 1.0, lead + (res, res, self.num_image_channels))
                    d["global_voltages"] = _box(-1.0, 1.0, lead + (n_glob,))
                obs[aid] = _dict(**d)
            else:
                obs[aid] = img
            act[aid] = _box(-1.0, 1.0, lead + (1,))
        self.observation_spaces = _dict(**obs)
        self.action_spaces = _dict(**act)
        self.observation_space = self.observation_spaces
        self.action_space = self.action_spaces

    # ---- observation extraction (:311-383), batched ------------------------------------------------------------
    def _extract_agent_observation(self, global_obs: dict, agent_id: str):
        """``global_obs['image']``: tensor / array ``[E, N-1, H, W]`` (the batched env's layout, channels first).  Returns this
        agent's observation with a leading env axis, channels LAST like the reference: ``image [E, H, W, C]``."""
        import torch
        image = global_obs["image"]
        if not torch.is_tensor(image):
            image = torch.as_tensor(np.asarray(image))
        kind, idx = agent_id.split("_")
        idx = int(idx)
        ch = self.agent_channel_map[agent_id]
        if len(ch) == 2:
            img1, img2 = image[:, ch[0]], image[:, ch[1]]                    # [E, H, W]
            if idx == self.num_gates - 1:                                      # final plunger: both transposed
                img1, img2 = img1.transpose(-1, -2), img2.transpose(-1, -2)
            elif idx != 0:                                                     # middle plungers: second transposed
                img2 = img2.transpose(-1, -2)
            agent_image = torch.stack([img1, img2], dim=-1)
        else:
            agent_image = image[:, ch[0]:ch[0] + 1].permute(0, 2, 3, 1)
        agent_image = agent_image.to(torch.float32)
        if not self.return_voltage:
            return self._leaf(agent_image)
        src = global_obs["obs_gate_voltages"] if kind == "plunger" else global_obs["obs_barrier_voltages"]
        out = {"image": self._leaf(agent_image), "voltage": np.asarray(src, dtype=np.float32)[:, idx:idx + 1]}
        if self.return_global_state:
            out["global_image"] = self._leaf(image.permute(0, 2, 3, 1).to(torch.float32))
            out["global_voltages"] = np.concatenate([np.asarray(global_obs["obs_gate_voltages"], dtype=np.float32),
                                                     np.asarray(global_obs["obs_barrier_voltages"], dtype=np.float32)], axis=1)
        return out

    def _leaf(self, t):
        return t.contiguous().cpu().numpy() if self.to_numpy else t

    def _squeeze(self, tree):
        """n_env == 1: drop the env axis so that every leaf has the reference's shape."""
        if self.n_env != 1:
            return tree
        if isinstance(tree, dict):
            return {k: self._squeeze(v) for k, v in tree.items()}
        return tree[0]

    def _observations(self, global_obs):
        if global_obs is None:
            return None
        return {aid: self._squeeze(self._extract_agent_observation(global_obs, aid)) for aid in self.all_agent_ids}

    # ---- MultiAgentEnv API ---------------------------------------------------------------------------------------
    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.base_env.rng = np.random.default_rng(seed)
            self.base_env.seed = int(seed)
        global_obs, global_info = self.base_env.reset()
        self._ma._returns = np.zeros(self.n_env)
        self._ma._lengths = np.zeros(self.n_env, dtype=np.int64)
        self._ma._last_info = global_info
        return self._observations(global_obs), {aid: global_info for aid in self.all_agent_ids}

    def step(self, action_dict: dict):
        assert len(action_dict) == len(self.all_agent_ids), "Agent actions must match the number of agents"
        assert all(aid in self._agent_ids for aid in action_dict), "Unknown agent IDs in actions"
        acts = {aid: np.asarray(a, dtype=np.float32).reshape(self.n_env, -1) for aid, a in action_dict.items()}
        gate, barrier = self._ma._combine_agent_actions(acts)
        global_obs, rewards, terminated, truncated, info = self.base_env.step(gate, barrier)
        self._ma._returns += rewards["gates"].sum(axis=1) + rewards["barriers"].sum(axis=1)
        self._ma._lengths += 1
        self._ma._last_info = info
        dist = self._ma._distribute_rewards(rewards)
        infos = self._ma._infos(info)
        if self.n_env == 1:
            rew = {aid: float(r[0]) for aid, r in dist.items()}
            term = {aid: bool(terminated[0]) for aid in self.all_agent_ids}
            trunc = {aid: bool(truncated[0]) for aid in self.all_agent_ids}
            infos = {aid: {k: v[0] for k, v in d.items()} for aid, d in infos.items()}
        else:
            rew = dist
            term = {aid: terminated for aid in self.all_agent_ids}
            trunc = {aid: truncated for aid in self.all_agent_ids}
        term["__all__"] = bool(np.all(terminated))
        trunc["__all__"] = bool(np.all(truncated))
        return self._observations(global_obs), rew, term, trunc, infos

    # ---- vector helpers ------------------------------------------------------------------------------------------
    def sub_env(self, tree, e: int):
        """Slice env ``e`` out of a batched return value (observation / reward / info dict): the reference's shapes."""
        if isinstance(tree, dict):
            return {k: self.sub_env(v, e) for k, v in tree.items()}
        if isinstance(tree, (bool, float, int)) or tree is None:
            return tree
        return tree[e]

    def episode_stats(self):
        return self._ma.episode_stats()

    def close(self):
        pass


def env_creator(env_config: dict | None = None):
    """``register_env('qarray_multiagent_env', env_creator)``-compatible factory (training/train.py:532-539): keys ``n_env``,
    ``num_dots``, ``device``, ``seed``, ``return_voltage``, ``return_global_state`` and any ``EnvConfig`` field."""
    from .engine import Engine
    cfg = dict(env_config or {})
    n_env, num_dots = int(cfg.pop("n_env", 1)), int(cfg.pop("num_dots", 4))
    device, seed = int(cfg.pop("device", 0)), int(cfg.pop("seed", 0))
    rv, rg = bool(cfg.pop("return_voltage", True)), bool(cfg.pop("return_global_state", False))
    cap = cfg.pop("capacitance_model", None)
    fields = {k: v for k, v in cfg.items() if k in EnvConfig.__dataclass_fields__}
    return VectorMultiAgentEnv(n_env, num_dots, engine=Engine(device), config=EnvConfig(**fields), seed=seed,
                               capacitance_model=cap, return_voltage=rv, return_global_state=rg)
