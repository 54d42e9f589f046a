import sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "rl-agent-for-qubit-array-tuning_b200"), os.path.join(os.getcwd(), "tools")]
import numpy as np, torch
import sweep
from qdsim import Engine, synth, FLAG_THERMAL, N_F32, N_U8
eng = Engine(0)
dev = synth.sample_devices(2048, 8, seed=8)
mbt = synth.model_batch(dev, thermal=True)
sct = synth.env_step_scans(mbt, dev, res=64, seed=9)
sweep.FLAGS = sweep.FLAGS | FLAG_THERMAL
print("RESULT thermal", sweep.time_gpu(eng, mbt, sct, N_F32))
