"""Per-agent views of a batched observation (rank 3 of SURVEY.md section 8f): the channel assignment and transposes of
``MultiAgentEnvWrapper`` (src/qadapt/environment/multi_agent_wrapper.py:147-178, 311-383) for a whole env batch, as
zero-copy torch views of the ``[env, pair, iy, ix]`` image produced by ``qdsim.obs.observe`` / ``BatchedDeviceEnv``.

Reference rule (per env, image (H, W, N-1)):
  plunger_0     -> channels [0, 0], no transpose
  plunger_i     -> channels [i-1, i], the second transposed            (0 < i < N-1)
  plunger_{N-1} -> channels [N-2, N-2], both transposed
  barrier_j     -> channel  [j]
Agent images here are ``[env, 2 or 1, H, W]`` (channels first); the reference's are ``(H, W, 2 or 1)`` per env.
"""
from __future__ import annotations


def agent_ids(num_dots: int):
    return [f"plunger_{i}" for i in range(num_dots)] + [f"barrier_{j}" for j in range(num_dots - 1)]


def agent_image(image, agent_id: str):
    """``image``: torch tensor ``[E, N-1, H, W]`` -> this agent's view ``[E, 2, H, W]`` (plunger) or ``[E, 1, H, W]``."""
    import torch
    n_pairs = image.shape[1]
    kind, idx = agent_id.split("_")
    idx = int(idx)
    if kind == "barrier":
        return image[:, idx:idx + 1]
    if idx == 0:
        a = b = image[:, 0]
    elif idx == n_pairs:                      # last plunger (N-1): last channel twice, both transposed
        a = b = image[:, n_pairs - 1].transpose(-1, -2)
    else:
        a, b = image[:, idx - 1], image[:, idx].transpose(-1, -2)
    return torch.stack([a, b], dim=1)


def agent_observations(obs: dict, num_dots: int, return_voltage: bool = True):
    """Dict agent_id -> {'image': [E, c, H, W], 'voltage': [E, 1]} (or the image alone), from a batched observation."""
    out = {}
    for aid in agent_ids(num_dots):
        img = agent_image(obs["image"], aid)
        if not return_voltage:
            out[aid] = img
            continue
        kind, idx = aid.split("_")
        v = obs["obs_gate_voltages"] if kind == "plunger" else obs["obs_barrier_voltages"]
        out[aid] = {"image": img, "voltage": v[:, int(idx):int(idx) + 1]}
    return out
