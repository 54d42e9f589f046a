"""ctypes wrapper of the plain-C restatement (oracle/cport/qd_cport.c).  TEST INFRASTRUCTURE (oracle/__init__.py):
used by tests/ to cross-check the NumPy oracle and by bench.py as the CPU baseline / ``--impl reference`` arm."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "qd_cport.c")
_LIB = os.path.join(_HERE, "libqd_cport.so")
_lib = None

ALG = {"default": 0, "thresholded": 1, "brute_force": 2}


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-o", _LIB, _SRC, "-lm"])
    return _LIB


def available() -> bool:
    try:
        _load()
        return True
    except (OSError, subprocess.CalledProcessError):
        return False


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        lib = C.CDLL(_LIB)
        vp = C.c_void_p
        lib.qd_cport_scans.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_uint, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
        lib.qd_cport_scans.restype = C.c_int
        _lib = lib
    return _lib


def run_scans(mb, scans, flags: int, threads: int = 0, want_n: bool = True, want_margin: bool = False):
    """``mb``: a ModelBatch-like object (cdd_inv_gs, cdd_gs, cdd_inv_full, cgd_full, params, algorithm);
    ``scans``: qd_scan records.  Returns (z float32 [pixels], n float64 [pixels, N] or None, seconds) -- with
    ``want_margin`` a fourth item: the best / second-best candidate energy gap of every pixel, float64 [pixels]."""
    lib = _load()
    scans = np.ascontiguousarray(scans)
    pixels = int((scans["pix_offset"] + scans["nx"].astype(np.int64) * scans["ny"]).max())
    n_dot = mb.cdd_inv_gs.shape[-1]
    z = np.empty(pixels, dtype=np.float32)
    n = np.empty((pixels, n_dot), dtype=np.float64) if want_n else None
    margin = np.full(pixels, np.inf) if want_margin else None
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (mb.cdd_inv_gs, mb.cdd_gs, mb.cdd_inv_full, mb.cgd_full)]
    params = np.ascontiguousarray(mb.params)
    t0 = time.perf_counter()
    rc = lib.qd_cport_scans(len(scans), scans.ctypes.data, n_dot, mb.cgd_full.shape[-1], ALG[mb.algorithm], flags,
                            *[a.ctypes.data for a in arrs], params.ctypes.data, z.ctypes.data,
                            None if n is None else n.ctypes.data, None if margin is None else margin.ctypes.data, threads)
    dt = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"qd_cport_scans failed: {rc}")
    if want_margin:
        return z, n, dt, margin
    return z, n, dt


def time_scans(mb, scans, flags: int, threads: int = 0):
    """(pixels/s, pixels, seconds) of the C restatement over ``scans``."""
    scans = np.ascontiguousarray(scans).copy()
    scans["pix_offset"] = np.arange(len(scans), dtype=np.int64) * (scans["nx"].astype(np.int64) * scans["ny"])
    _, _, dt = run_scans(mb, scans, flags, threads=threads, want_n=False)
    pixels = int((scans["nx"].astype(np.int64) * scans["ny"]).sum())
    return pixels / dt, pixels, dt


# ---- Path B (tunnel-coupled ground state): oracle/cport/qd_cport_b.c ----------------------------------------------------
_SRC_B = os.path.join(_HERE, "qd_cport_b.c")
_LIB_B = os.path.join(_HERE, "libqd_cport_b.so")
_lib_b = None


def _load_b():
    global _lib_b
    if _lib_b is None:
        if not os.path.exists(_LIB_B) or os.path.getmtime(_LIB_B) < os.path.getmtime(_SRC_B):
            subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-o", _LIB_B, _SRC_B, "-lm"])
        lib = C.CDLL(_LIB_B)
        vp = C.c_void_p
        lib.qd_cport_tunnel_points.argtypes = [C.c_long, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_double, vp, C.c_double,
                                               C.c_double, vp, vp, vp, C.c_int]
        lib.qd_cport_tunnel_points.restype = C.c_int
        _lib_b = lib
    return _lib_b


def tunnel_ground_state(m, v_ext, threads: int = 1):
    """``m``: oracle.scan.Model (algorithm "tunnel"); ``v_ext`` (P, n_volt).  Returns (<n> (P, N), gap (P,), seconds):
    the reference's formulation in C (all 4^N candidates, 32 x 32 Jacobi), OpenMP over pixels."""
    lib = _load_b()
    v = np.ascontiguousarray(v_ext, dtype=np.float64)
    cinv = np.ascontiguousarray(m.cdd_inv, dtype=np.float64)
    a = np.ascontiguousarray(m.cgd, dtype=np.float64)
    n, nv = cinv.shape[0], a.shape[1]
    cbg = None if m.cbg is None or nv <= m.n_gate else np.ascontiguousarray(m.cbg, dtype=np.float64)
    alpha = np.ascontiguousarray(np.asarray(m.alpha, dtype=np.float64) if m.alpha is not None else np.zeros(8))
    out = np.empty((v.shape[0], n))
    gap = np.empty(v.shape[0])
    t0 = time.perf_counter()
    rc = lib.qd_cport_tunnel_points(v.shape[0], n, nv, m.n_gate, cinv.ctypes.data, a.ctypes.data,
                                    None if cbg is None else cbg.ctypes.data, float(m.tc_base), alpha.ctypes.data,
                                    float(getattr(m, "vc_alpha", 0.0)), float(getattr(m, "vc_beta", 0.0)), v.ctypes.data,
                                    out.ctypes.data, gap.ctypes.data, int(threads))
    dt = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"qd_cport_tunnel_points failed: {rc}")
    return out, gap, dt
