#!/usr/bin/env python
"""BASELINE config 1: latency of ONE 2-dot 64 x 64 do2d_open -- through the drop-in Python class, through the bare C ABI call,
and the C restatement on 1 and on all host cores.  QDSIM_TRACE=1 adds the library's own phase breakdown on stderr."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rl-agent-for-qubit-array-tuning_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    import qarray
    from qdsim import N_U8
    from qdsim.engine import new_scans
    from qdsim.runtime import engine_for, shutdown
    m = qarray.ChargeSensedDotArray(Cdd=[[0, .12], [.12, 0]], Cgd=[[1.0, .35, 0], [.3, .97, 0]], Cds=[[.04, .045]],
                                    Cgs=[[6e-5, 3e-5, .98]], coulomb_peak_width=0.15, T=0.0, algorithm="default",
                                    implementation="jax", max_charge_carriers=4)
    args = (1, -3.3, 0.7, 64, 2, -3.1, 0.9, 64)
    for _ in range(200):
        m.do2d_open(*args)
    reps = 2000
    t0 = time.perf_counter()
    for _ in range(reps):
        m.do2d_open(*args)
    lat_class = (time.perf_counter() - t0) / reps
    eng = engine_for(m)
    s = new_scans(1)
    v0, dx, dy = m.gate_voltage_composer.affine2d(*args)
    s["v0"][0, :3], s["dx"][0, :3], s["dy"][0, :3], s["nx"], s["ny"], s["peak_width"] = v0, dx, dy, 64, 64, 0.15
    z = np.empty(4096, dtype=np.float32)
    n = np.empty((4096, 2), dtype=np.uint8)
    fn, ctx = eng._lib.qd_scan_open_host, eng._ctx
    sp, zp, nn = s.ctypes.data, z.ctypes.data, n.ctypes.data
    for _ in range(200):
        fn(ctx, 1, sp, zp, nn, N_U8, 0)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(ctx, 1, sp, zp, nn, N_U8, 0)
    lat_c = (time.perf_counter() - t0) / reps
    import bench
    mb1 = m._model_batch()
    out = {"do2d_open_python_class_us": lat_class * 1e6, "qd_scan_open_host_ctypes_us": lat_c * 1e6,
           "cpu_cport_all_cores_us": bench.cpu_reference_scan_seconds(mb1, s, 0, os.cpu_count()) * 1e6,
           "cpu_cport_1_core_us": bench.cpu_reference_scan_seconds(mb1, s, 0, 1) * 1e6, "cores": os.cpu_count()}
    print(json.dumps(out))
    shutdown()


if __name__ == "__main__":
    main()
