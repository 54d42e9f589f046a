#!/usr/bin/env python
"""BASELINE.json configs 2, 3 and 5 in one run (config 1 is a parity test, config 4 is bench.py's headline):

  config 2: 4-dot, 1024 envs, 64x64, latching + noise           -- Path A (default) and Path B (tunnel-coupled)
  config 3: 6-dot, 4096 envs (--quick: 256), brute-force search  -- Path A brute_force, max_charge_carriers = 4
  config 5: N in {2,4,6,8} x res in {32,64,128,256}, Path A default vs the CPU restatement (C port, all cores)

Device-resident timing (descriptors + models in HBM), CUDA events, 3 warm-up launches.  Writes one JSON document.
    python tools/sweep.py [--quick] > profiles/rNN_sweep.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rl-agent-for-qubit-array-tuning_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from qdsim import FLAG_LATCH, FLAG_NOISE, FLAG_RADIAL, N_F32, N_U8, Engine, synth  # noqa: E402

FLAGS = FLAG_LATCH | FLAG_NOISE | FLAG_RADIAL


def time_gpu(eng, mb, scans, n_type, steps=3, warmup=3):
    eng.set_models(mb)
    pixels = int(scans["nx"].astype(np.int64) @ scans["ny"])
    z = torch.empty(pixels, dtype=torch.float32, device="cuda")
    n = torch.empty((pixels, mb.n_dot), dtype=torch.uint8 if n_type == N_U8 else torch.float32, device="cuda")
    st = torch.cuda.current_stream()
    eng.scan_upload(scans, st)
    for _ in range(warmup):
        eng.scan_launch(z, n, n_type, FLAGS, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(steps):
        eng.scan_launch(z, n, n_type, FLAGS, st)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"pixels": pixels, "ms_per_step": ms, "pixels_per_s": pixels / (ms * 1e-3),
            "env_steps_per_s": mb.n_env / (ms * 1e-3)}


def time_cpu(mb, scans, budget_s=3.0):
    """CPU comparator = bench.py's cpu_baseline leg (the only place besides tests/ that may execute oracle/)."""
    import bench
    cores = os.cpu_count() or 1
    pps, pixels, dt, _ = bench.cpu_reference_pixels_per_s(mb, scans, FLAGS, 0, cores, budget_s=budget_s)
    return {"pixels_per_s": pps, "cores": cores, "sample_scans": pixels // int(scans["nx"][0] * scans["ny"][0]), "seconds": dt}


def next_rows(eng, out, quick=False):
    """batched env.step with CNN + Kalman virtualisation in the loop (BASELINE config 3's
    'Kalman virtualisation in the loop'), and dataset generation, wall-clock around whole calls."""
    import time
    from qdsim.dataset import generate_batch
    from qdsim.vector_env import BatchedDeviceEnv, EnvConfig
    from qdsim.virtualisation import make_capacitance_cnn
    torch.manual_seed(0)
    cnn = make_capacitance_cnn(3).cuda().eval()
    rows = {}
    for n_dot, n_env in ((4, 1024), (6, 256 if quick else 1024)):
        for method in (None, "kalman"):
            env = BatchedDeviceEnv(n_env, n_dot, engine=eng, seed=1, capacitance_model=cnn if method else None,
                                   config=EnvConfig(resolution=64, max_steps=50, update_method=method))
            if env.vg_updater is not None:
                env.vg_updater.autocast_dtype = torch.bfloat16
            env.reset()
            rng = np.random.default_rng(0)
            acts = [(rng.uniform(-0.1, 0.1, (n_env, n_dot)), rng.uniform(-0.1, 0.1, (n_env, n_dot - 1))) for _ in range(4)]
            env.step(*acts[0])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for a in acts[1:]:
                env.step(*a)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            rows[f"{n_dot}dot_{n_env}env_64x64_tunnel_update_{method}"] = {"ms_per_step": dt * 1e3, "env_steps_per_s": n_env / dt}
            print(n_dot, n_env, method, dt, file=sys.stderr)
    out["batched_env_shell_virtualisation_in_loop"] = rows
    ds = {}
    for use_barriers, n_s in ((False, 4096), (True, 256 if quick else 1024)):
        generate_batch(eng, 64, 4, seed=1, use_barriers=use_barriers, res=100)
        t0 = time.perf_counter()
        generate_batch(eng, n_s, 4, seed=2, use_barriers=use_barriers, res=100)
        dt = time.perf_counter() - t0
        ds[f"4dot_100x100_{'tunnel' if use_barriers else 'path_A'}"] = {"samples": n_s, "samples_per_s": n_s / dt,
                                                                        "note": "images copied to host, reference layout"}
    out["dataset_generation"] = ds


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--next-rows-only", action="store_true",
                    help="only the SURVEY 8f rows: env shell with Kalman virtualisation in the loop, dataset generation")
    args = ap.parse_args()
    if args.next_rows_only:
        eng = Engine(0)
        out = {"gpu": torch.cuda.get_device_name(0)}
        next_rows(eng, out, args.quick)
        print(json.dumps(out, indent=1))
        return
    eng = Engine(0)
    out = {"gpu": torch.cuda.get_device_name(0), "flags": "latching + white/telegraph/radial noise, T=0"}

    # ---- config 1: single-scan latency through the drop-in class ----
    import time
    import qarray
    m = qarray.ChargeSensedDotArray(Cdd=[[0, .12], [.12, 0]], Cgd=[[1.0, .35, 0], [.3, .97, 0]], Cds=[[.04, .045]],
                                    Cgs=[[6e-5, 3e-5, .98]], coulomb_peak_width=0.15, T=0.0, algorithm="default",
                                    implementation="jax", max_charge_carriers=4)
    for _ in range(20):
        m.do2d_open(1, -3.3, 0.7, 64, 2, -3.1, 0.9, 64)
    t0 = time.perf_counter()
    for _ in range(200):
        m.do2d_open(1, -3.3, 0.7, 64, 2, -3.1, 0.9, 64)
    lat = (time.perf_counter() - t0) / 200
    import bench
    from qdsim.engine import new_scans
    mb1 = m._model_batch()
    s1 = new_scans(1)
    v0, dx, dy = m.gate_voltage_composer.affine2d(1, -3.3, 0.7, 64, 2, -3.1, 0.9, 64)
    s1["v0"][0, :3], s1["dx"][0, :3], s1["dy"][0, :3], s1["nx"], s1["ny"], s1["peak_width"] = v0, dx, dy, 64, 64, 0.15
    dt_cpu = bench.cpu_reference_scan_seconds(mb1, s1, 0, os.cpu_count())
    dt_cpu1 = bench.cpu_reference_scan_seconds(mb1, s1, 0, 1)
    out["config1_2dot_64x64_single_do2d_open"] = {
        "gpu_call_latency_us": lat * 1e6, "note": "Python do2d_open -> qd_scan_open_host, host buffers, synchronous",
        "cpu_cport_all_cores_us": dt_cpu * 1e6, "cpu_cport_1_core_us": dt_cpu1 * 1e6}

    # ---- config 4 variant: thermal (T ~ U[50, 200] mK as the reference samples it), non-integer occupations ----
    from qdsim import FLAG_THERMAL
    global FLAGS
    dev = synth.sample_devices(2048, 8, seed=8)
    mbt = synth.model_batch(dev, thermal=True)
    sct = synth.env_step_scans(mbt, dev, res=64, seed=9)
    keep = FLAGS
    FLAGS = FLAGS | FLAG_THERMAL
    out["config4_variant_8dot_2048env_thermal"] = time_gpu(eng, mbt, sct, N_F32)
    FLAGS = keep

    # ---- batched env shell (first "next" row): whole env.step incl. host logic, scans, device normalisation ----
    from qdsim.vector_env import BatchedDeviceEnv, EnvConfig
    shell = {}
    for n_dot, n_env in ((4, 1024), (8, 256)):
        env = BatchedDeviceEnv(n_env, n_dot, engine=eng, config=EnvConfig(resolution=64, max_steps=50), seed=1)
        env.reset()
        rng = np.random.default_rng(0)
        acts = [(rng.uniform(-0.1, 0.1, (n_env, n_dot)), rng.uniform(-0.1, 0.1, (n_env, n_dot - 1))) for _ in range(4)]
        env.step(*acts[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for a in acts[1:]:
            env.step(*a)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        shell[f"{n_dot}dot_{n_env}env_64x64_tunnel"] = {"ms_per_step": dt * 1e3, "env_steps_per_s": n_env / dt}
    out["batched_env_shell"] = shell

    # ---- config 2 ----
    dev = synth.sample_devices(1024, 4, seed=2)
    mb = synth.model_batch(dev)
    sc = synth.env_step_scans(mb, dev, res=64, seed=3)
    c2 = {"path_A_default": time_gpu(eng, mb, sc, N_U8), "cpu_path_A": time_cpu(mb, sc)}
    devb = synth.sample_barrier_devices(1024, 4, seed=2)
    mbb = synth.tunnel_batch(devb)
    scb = synth.env_step_scans(mbb, devb, res=64, seed=3)
    c2["path_B_tunnel"] = time_gpu(eng, mbb, scb, N_F32, steps=2, warmup=1)
    out["config2_4dot_1024env_64x64"] = c2

    # ---- config 3 ----
    n_env3 = 256 if args.quick else 4096
    dev = synth.sample_devices(n_env3, 6, seed=4)
    mb = synth.model_batch(dev, algorithm="brute_force", max_charge_carriers=4)
    sc = synth.env_step_scans(mb, dev, res=64, seed=5)
    c3 = {"path_A_brute_force": time_gpu(eng, mb, sc, N_U8, steps=2, warmup=1)}
    c3["cpu_brute_force"] = time_cpu(mb, sc[:64], budget_s=3.0)
    mbd = synth.model_batch(dev, algorithm="default")
    c3["path_A_default_same_devices"] = time_gpu(eng, mbd, sc, N_U8)
    # ... and as BASELINE states it: brute-force search AND Kalman virtualisation in the loop -- the batched env shell on
    # Path A brute_force models, CNN forward + scalar Kalman filter + batched VGM update after every observation
    from qdsim.vector_env import BatchedDeviceEnv, EnvConfig
    from qdsim.virtualisation import make_capacitance_cnn
    torch.manual_seed(0)
    cnn = make_capacitance_cnn(3).cuda().eval()
    for method in (None, "kalman"):
        env = BatchedDeviceEnv(n_env3, 6, engine=eng, seed=1, capacitance_model=cnn if method else None,
                               config=EnvConfig(resolution=64, max_steps=50, update_method=method, use_barriers=False,
                                                algorithm="brute_force", max_charge_carriers=4))
        if env.vg_updater is not None:
            env.vg_updater.autocast_dtype = torch.bfloat16
        env.reset()
        rng = np.random.default_rng(0)
        acts = [(rng.uniform(-0.1, 0.1, (n_env3, 6)), rng.uniform(-0.1, 0.1, (n_env3, 5))) for _ in range(4)]
        env.step(*acts[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for a in acts[1:]:
            env.step(*a)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        c3[f"env_shell_brute_force_update_{method}"] = {"ms_per_step": dt * 1e3, "env_steps_per_s": n_env3 / dt,
                                                         "pixels_per_s": n_env3 * 5 * 64 * 64 / dt}
    out[f"config3_6dot_{n_env3}env_64x64"] = c3

    # ---- config 5 ----
    sweep = []
    for n_dot in (2, 4, 6, 8):
        for res in (32, 64, 128, 256):
            n_env = max(16, int((2e8 if not args.quick else 2e7) / ((n_dot - 1) * res * res)))
            dev = synth.sample_devices(n_env, n_dot, seed=6)
            mb = synth.model_batch(dev)
            sc = synth.env_step_scans(mb, dev, res=res, seed=7)
            g = time_gpu(eng, mb, sc, N_U8)
            c = time_cpu(mb, sc, budget_s=1.5)
            sweep.append({"n_dot": n_dot, "res": res, "n_env": n_env, "gpu_pixels_per_s": g["pixels_per_s"],
                          "gpu_ms": g["ms_per_step"], "cpu_pixels_per_s": c["pixels_per_s"], "cpu_cores": c["cores"],
                          "ratio": g["pixels_per_s"] / c["pixels_per_s"]})
            print(sweep[-1], file=sys.stderr)
    out["config5_sweep_path_A_default"] = sweep
    out["launches"] = eng.launch_count
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
