import sys, time, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "rl-agent-for-qubit-array-tuning_b200")]
import numpy as np, torch
from qdsim import Engine, obs as qobs, FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL
from qdsim.vector_env import BatchedDeviceEnv, EnvConfig
eng = Engine(0)
for n_dot, n_env in ((4, 1024), (8, 256)):
    env = BatchedDeviceEnv(n_env, n_dot, engine=eng, config=EnvConfig(resolution=64), seed=1)
    env.reset()
    rng = np.random.default_rng(0)
    env.step(rng.uniform(-.1, .1, (n_env, n_dot)), rng.uniform(-.1, .1, (n_env, n_dot - 1)))
    torch.cuda.synchronize()
    t0 = time.perf_counter(); scans = env._scans(); t1 = time.perf_counter()
    flags = FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL
    st = torch.cuda.current_stream()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(st)
    eng.scan_open(scans, env.z_dev, None, 0, flags, st)
    e[1].record(st)
    eng.normalise_obs(env.z_dev, stream=st)
    e[2].record(st)
    torch.cuda.synchronize()
    t2 = time.perf_counter(); env._ground_truth(); env._reward(); t3 = time.perf_counter()
    print(n_dot, n_env, "host scans %.1f ms | gpu scan %.1f ms | normalise %.1f ms | gt+reward %.1f ms | replaced scans %.2f" % (
        (t1 - t0) * 1e3, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), (t3 - t2) * 1e3, (scans["rad_mode"] == 2).mean()))
