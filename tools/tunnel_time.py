#!/usr/bin/env python
"""Device-resident timing of the tunnel-coupled path (Path B) at several sizes, one JSON line each; environment switches
(QDSIM_LIB, QDSIM_TUNNEL_MONO, QDSIM_TUNNEL_CHUNK_PIX) are read by the library, so one gpurun call can A/B variants:

    python tools/tunnel_time.py [--cases 4:1024,6:512,8:512] [--steps 3] [--tag name]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rl-agent-for-qubit-array-tuning_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from qdsim import FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL, N_F32, Engine, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="4:1024,6:512,8:512")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    eng = Engine(0)
    flags = FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL
    for case in a.cases.split(","):
        n_dot, n_env = map(int, case.split(":"))
        dev = synth.sample_barrier_devices(n_env, n_dot, seed=1234)
        mb = synth.tunnel_batch(dev)
        eng.set_models(mb)
        scans = synth.env_step_scans(mb, dev, res=a.res, seed=99)
        pixels = len(scans) * a.res * a.res
        z = torch.empty(pixels, dtype=torch.float32, device="cuda")
        n = torch.empty((pixels, n_dot), dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream()
        eng.scan_upload(scans, st)
        for _ in range(2):
            eng.scan_launch(z, n, N_F32, flags, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(a.steps):
            eng.scan_launch(z, n, N_F32, flags, st)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print(json.dumps({"tag": a.tag, "mono": os.environ.get("QDSIM_TUNNEL_MONO", "0"), "lib": os.environ.get("QDSIM_LIB", ""),
                          "n_dot": n_dot, "n_env": n_env, "pixels": pixels, "ms": ms, "Mpix_s": pixels / ms / 1e3,
                          "env_steps_s": n_env / ms * 1e3, "nsum": float(n.double().sum().item()),
                          "nan_pixels": int(torch.isnan(n).any(dim=1).sum().item()), "opt": os.environ.get("QDSIM_TUNNEL_OPT", ""),
                          "first_nan": (torch.isnan(n).any(dim=1).nonzero()[:4, 0].tolist())}), flush=True)


if __name__ == "__main__":
    main()
