import sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "rl-agent-for-qubit-array-tuning_b200"))
import torch
from qdsim import Engine, synth, FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL, N_U8
eng = Engine(0)
dev = synth.sample_devices(4096, 8, seed=1234); mb = synth.model_batch(dev); eng.set_models(mb)
sc = synth.env_step_scans(mb, dev, res=64, seed=99)
pix = len(sc) * 4096
z = torch.empty(pix, dtype=torch.float32, device="cuda"); n = torch.empty((pix, 8), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream(); eng.scan_upload(sc, st)
def t(flags, ntype=N_U8, zz=z):
    for _ in range(2): eng.scan_launch(zz, n, ntype, flags, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(3): eng.scan_launch(zz, n, ntype, flags, st)
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3
for name, fl in (("none", 0), ("latch", FLAG_LATCH), ("noise", FLAG_NOISE), ("radial", FLAG_RADIAL), ("latch+noise", FLAG_LATCH | FLAG_NOISE), ("all", FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL)):
    print(f"{name:12s} {t(fl):7.2f} ms   (4096 envs; x4 for 16384)")
