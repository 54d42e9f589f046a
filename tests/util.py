"""Shared helpers of the parity tests: turn the product-side batch (ModelBatch + qd_scan records) into the oracle's
Model / Scan objects so both sides see the same numbers."""
from __future__ import annotations

import numpy as np

from oracle import scan as oscan


def oracle_model(mb, e: int, flags: int = 0):
    from qdsim import FLAG_LATCH, FLAG_NOISE, FLAG_THERMAL
    p = mb.params[e]
    n = mb.n_dot
    noise = bool(flags & FLAG_NOISE)
    return oscan.Model(
        cdd_inv=mb.cdd_inv_gs[e], cdd=None if mb.cdd_gs is None else mb.cdd_gs[e],
        cgd=mb.cgd_full[e, :n, :], cdd_inv_full=mb.cdd_inv_full[e], cgd_full=mb.cgd_full[e],
        algorithm=mb.algorithm, threshold=float(p["threshold"]), max_charge_carriers=int(p["max_charge_carriers"]),
        kT=float(p["kT"]) if flags & FLAG_THERMAL else 0.0,
        latching=bool(p["latching"]) and bool(flags & FLAG_LATCH),
        p_leads=p["p_leads"][:n].copy(), p_inter=p["p_inter"].reshape(8, 8)[:n, :n].copy(),
        white_amp=float(p["white_amp"]) if noise else 0.0, tele_p01=float(p["tele_p01"]),
        tele_p10=float(p["tele_p10"]), tele_amp=float(p["tele_amp"]) if noise else 0.0,
        n_gate=mb.n_gate, cbg=None if mb.cbg is None else mb.cbg[e], tc_base=float(p["tc_base"]),
        alpha=p["alpha"].copy(), num_charge_states=mb.num_charge_states,
        charge_state_batch_size=mb.charge_state_batch_size, vc_alpha=float(p["vc_alpha"]), vc_beta=float(p["vc_beta"]))


def oracle_scan(rec, n_volt: int, flags: int = 0):
    from qdsim import FLAG_RADIAL
    rad_mode = int(rec["rad_mode"]) if flags & FLAG_RADIAL else 0
    return oscan.Scan(
        v0=rec["v0"][:n_volt].copy(), dx=rec["dx"][:n_volt].copy(), dy=rec["dy"][:n_volt].copy(),
        nx=int(rec["nx"]), ny=int(rec["ny"]), peak_width=float(rec["peak_width"]), seed=int(rec["seed"]),
        rad_mode=rad_mode,
        rad=(float(rec["rad_x0"]), float(rec["rad_dx"]), float(rec["rad_y0"]), float(rec["rad_dy"]),
             float(rec["rad_alpha"]), float(rec["rad_zero_radius"]), float(rec["rad_max_amp"])))


def oracle_batch(mb, scans, flags: int = 0, which=None):
    """Run the oracle over ``scans[which]``; returns (z, n, margin) stacked, each scan (ny, nx[, N])."""
    from qdsim import FLAG_CARRY_ROWS, FLAG_LATCH_EXACT, FLAG_WHITE_ON_OUTPUT
    which = range(len(scans)) if which is None else which
    zs, ns, ms = [], [], []
    for i in which:
        rec = scans[i]
        m = oracle_model(mb, int(rec["env_id"]), flags)
        s = oracle_scan(rec, mb.n_volt, flags)
        z, n, mg = oscan.simulate_scan(
            m, s, latch_compare="exact" if flags & FLAG_LATCH_EXACT else "rounded",
            carry_rows=bool(flags & FLAG_CARRY_ROWS),
            white_on="output" if flags & FLAG_WHITE_ON_OUTPUT else "input", return_margin=True)
        zs.append(z)
        ns.append(n)
        ms.append(mg)
    return np.stack(zs), np.stack(ns), np.stack(ms)


def compare_charges(n_gpu, n_ref, margin, tie_tol=1e-9, max_tie_frac=0.005):
    """Bit-exact on every pixel whose best/second-best energy gap exceeds ``tie_tol`` (an exact tie has no defined
    winner across two summation orders); the near-tie set itself must be tiny."""
    n_gpu = np.asarray(n_gpu).astype(np.int64)
    n_ref = np.rint(np.asarray(n_ref)).astype(np.int64)
    safe = margin > tie_tol
    assert (~safe).mean() <= max_tie_frac, f"too many near-ties: {(~safe).mean():.4f}"
    bad = (n_gpu != n_ref).any(axis=-1) & safe
    assert not bad.any(), f"{bad.sum()} of {bad.size} pixels differ, first at {np.argwhere(bad)[:5]}"
    return int((~safe).sum())


def _params_from_bytes(raw):
    """Fixture bytes -> current PARAMS_DTYPE (fixtures made before ABI 2 lack the vc_alpha / vc_beta fields)."""
    from qdsim import PARAMS_DTYPE
    raw = np.ascontiguousarray(raw)
    if raw.size % PARAMS_DTYPE.itemsize == 0:
        return raw.view(PARAMS_DTYPE).copy()
    legacy = np.dtype([(n, PARAMS_DTYPE.fields[n][0]) for n in PARAMS_DTYPE.names if n not in ("vc_alpha", "vc_beta")],
                      align=True)
    assert legacy.itemsize == 712 and raw.size % 712 == 0, (legacy.itemsize, raw.size)
    old = raw.view(legacy)
    out = np.zeros(old.shape, dtype=PARAMS_DTYPE)
    for n in legacy.names:
        out[n] = old[n]
    return out


def load_golden(name):
    """tests/golden/<name>.npz -> (ModelBatch, scans, flags, z, n, margin): inputs rebuilt through the product's host code
    from the RAW capacitances stored in the fixture."""
    import os
    from qdsim import PARAMS_DTYPE, SCAN_DTYPE
    from qdsim.engine import ModelBatch
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")
    d = np.load(path)
    params = _params_from_bytes(d["params"])
    if str(d["algorithm"]) == "tunnel":
        from qdsim.engine import tunnel_model_batch
        mb = tunnel_model_batch(d["Cdd"], d["Cgd"], d["Cds"], d["Cgs"], d["Cbd"], d["Cbg"], d["Cbs"],
                                params["tc_base"], params["alpha"][:, :d["Cbd"].shape[-1]])
    else:
        mb = ModelBatch.from_capacitances(d["Cdd"], d["Cgd"], d["Cds"], d["Cgs"], algorithm=str(d["algorithm"]))
    mb.params = params
    scans = d["scans"].view(SCAN_DTYPE).copy()
    return mb, scans, int(d["flags"]), d["z"], d["n"], d["margin"]


GOLDEN_CASES = ["c1_2dot_64x64_noise_free", "c2_4dot_latched_full_noise", "c2b_4dot_flat_pass", "c3_6dot_brute_force",
                "c4_8dot_latched_full_noise", "t_3dot_thermal", "t_5dot_thresholded"]
GOLDEN_TUNNEL_CASES = ["b_4dot_tunnel_latched_noise", "b_6dot_tunnel_coupled"]
