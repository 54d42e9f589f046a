"""The N>1 host logic on CPU: env sharding and the episode-stat all-gather, world_size 2 and 3 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_env, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "rl-agent-for-qubit-array-tuning_b200"))
    from qdsim import parallel, synth
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(n_env, rank, world)
    # every rank builds only its own shard of devices and scan descriptors (what bench.py does per rank)
    dev = synth.sample_devices(hi - lo, 3, seed=100 + rank)
    mb = synth.model_batch(dev)
    scans = synth.env_step_scans(mb, dev, res=8, seed=1)
    assert len(scans) == (hi - lo) * 2 and scans["env_id"].max() == hi - lo - 1
    stats = torch.stack([torch.arange(lo, hi, dtype=torch.float32), torch.full((hi - lo,), float(rank))], dim=1)
    full = parallel.gather_episode_stats(stats, n_env)
    ok = full.shape == (n_env, 2) and torch.equal(full[:, 0], torch.arange(n_env, dtype=torch.float32))
    owners = [r for r in range(world) for _ in range(*parallel.shard_range(n_env, r, world))]
    ok = ok and torch.equal(full[:, 1], torch.tensor(owners, dtype=torch.float32))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def _rollout_worker(rank, world, port, n_env, out):
    """Each rank steps its own shard of envs (host logic only, no GPU) and the episode statistics meet on every rank."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "rl-agent-for-qubit-array-tuning_b200"))
    import numpy as np
    from qdsim import parallel
    from qdsim.multi_agent import BatchedMultiAgentEnv
    from qdsim.vector_env import BatchedDeviceEnv, EnvConfig
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(n_env, rank, world)
    env = BatchedMultiAgentEnv(BatchedDeviceEnv(hi - lo, 3, engine=None, config=EnvConfig(resolution=8, max_steps=2),
                                                seed=50 + rank))
    env.reset()
    rng = np.random.default_rng(rank)
    for _ in range(2):
        _, _, _, trunc, _ = env.step({a: rng.uniform(-1, 1, size=(hi - lo, 1)) for a in env.all_agent_ids}, skip_obs=True)
    local = torch.from_numpy(env.episode_stats())
    full = parallel.gather_episode_stats(local, n_env)
    ok = trunc["__all__"] and full.shape == (n_env, 4) and torch.equal(full[lo:hi], local) and bool((full[:, 1] == 2).all())
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_rollout_statistics_over_gloo():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_rollout_worker, args=(2, port, 7, out), nprocs=2, join=True)
        assert all(out[r] for r in range(2)) and len(out) == 2


@pytest.mark.parametrize("world,n_env", [(2, 10), (3, 10), (2, 7)])
def test_sharding_and_stat_gather_over_gloo(world, n_env):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_env, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world)) and len(out) == world


def test_shard_ranges_partition_the_envs():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "rl-agent-for-qubit-array-tuning_b200"))
    from qdsim import parallel
    for n_env in (1, 7, 16384, 16385):
        for world in (1, 2, 3, 8):
            ranges = [parallel.shard_range(n_env, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n_env
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 4, 4)


def test_numa_binding_is_a_no_op_without_topology():
    """No GPU / no sysfs entry: the helper reports False and leaves the affinity alone."""
    import os
    from qdsim import parallel
    before = os.sched_getaffinity(0)
    assert parallel.gpu_numa_cpus(0) is None or isinstance(parallel.gpu_numa_cpus(0), set)
    ok = parallel.bind_to_gpu_numa_node(0)
    assert ok or os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
