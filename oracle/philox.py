"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11, "Parallel random numbers: as easy as 1, 2, 3").

Test infrastructure (see ``oracle/__init__.py``).  The reference draws from the process-global ``np.random``
(``qarray_base_class.py:468, 492`` and the upstream noise / latching classes), which is neither seedable per scan nor
reproducible on a GPU.  The framework replaces it by this stream; the draw *distributions* are the reference's.

Stream contract shared with ``csrc/qd_philox.cuh``:

* key      = (seed & 0xffffffff, seed >> 32)                  -- one 64-bit seed per scan
* counter  = (index_lo, index_hi, purpose, 0)
* purpose 0: per-pixel block, index = iy*nx + ix.  words -> (w0, w1, w2, w3)
    - u1 = ((w0 >> 8) + 1) * 2^-24  in (0, 1],   u2 = (w1 >> 8) * 2^-24 in [0, 1)
    - z_white  = sqrt(-2 ln u1) * cos(2 pi u2)   (sensor input white noise,    unit normal)
    - z_radial = sqrt(-2 ln u1) * sin(2 pi u2)   (QADAPT radial / replacement, unit normal)
    - u_latch  = (w2 >> 8) * 2^-24               (latching acceptance draw)
    - u_tele   = (w3 >> 8) * 2^-24               (telegraph transition draw)
* purpose 1: per-row block, index = iy.  w0 -> u_row = (w0 >> 8) * 2^-24 (telegraph start state of a row when
  chains are per row).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)
_SHIFT32 = np.uint64(32)

INV_2_24 = 1.0 / 16777216.0


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds.  All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(*(np.asarray(a, dtype=np.uint32) for a in (c0, c1, c2, c3, k0, k1)))
    c0, c1, c2, c3, k0, k1 = (a.copy() for a in (c0, c1, c2, c3, k0, k1))
    with np.errstate(over="ignore"):
        for r in range(10):
            if r > 0:
                k0 = k0 + W0
                k1 = k1 + W1
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> _SHIFT32).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> _SHIFT32).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
    return c0, c1, c2, c3


def _key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32)


def words(seed: int, index, purpose: int = 0):
    """Raw 4x uint32 words for flattened indices ``index`` (any integer array) of stream ``purpose``."""
    index = np.asarray(index, dtype=np.uint64)
    k0, k1 = _key(seed)
    lo = (index & _MASK32).astype(np.uint32)
    hi = (index >> _SHIFT32).astype(np.uint32)
    return philox4x32_10(lo, hi, np.uint32(purpose), np.uint32(0), k0, k1)


def u24(w):
    """uint32 word -> float64 uniform in [0, 1) on a 2^-24 lattice (exactly representable in fp32 too)."""
    return (np.asarray(w, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) * INV_2_24


def pixel_draws(seed: int, n_pixels: int):
    """Per-pixel draws of a scan: dict(z_white, z_radial, u_latch, u_tele), each float64 of shape (n_pixels,)."""
    w0, w1, w2, w3 = words(seed, np.arange(n_pixels, dtype=np.uint64), 0)
    u1 = ((w0 >> np.uint32(8)).astype(np.float64) + 1.0) * INV_2_24
    u2 = u24(w1)
    r = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    return {
        "z_white": r * np.cos(ang),
        "z_radial": r * np.sin(ang),
        "u_latch": u24(w2),
        "u_tele": u24(w3),
    }


def row_draws(seed: int, n_rows: int):
    """Per-row uniform (telegraph start state when chains restart at each row)."""
    w0, _, _, _ = words(seed, np.arange(n_rows, dtype=np.uint64), 1)
    return u24(w0)
