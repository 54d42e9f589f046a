"""NumPy prototype of the eigen stage's Noda iteration (design study for qd_tunnel_eigen_kernel; not shipped code).

Per pixel: sectors of the 32-state basis by total charge; per sector A = Z-matrix (off-diagonals -|t|); Noda iteration
(inverse iteration with the Collatz-Wielandt lower bound as the shift, LDL^T without pivoting) on all live sectors;
sector exclusion by rigorous bounds; counts factorisations / solves and compares <n> with LAPACK.
"""
import sys
import numpy as np

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests")
sys.path.insert(0, "/root/repo/rl-agent-for-qubit-array-tuning_b200")
from oracle import path_b, composer  # noqa: E402


def ldl_solve(A, sigma, x, floor):
    m = len(x)
    M = A - sigma * np.eye(m)
    L = np.eye(m)
    d = np.zeros(m)
    M = M.copy()
    nfloor = 0
    for k in range(m):
        p = M[k, k]
        if not (p > floor):
            p = floor
            nfloor += 1
        d[k] = p
        l = M[k + 1:, k] / p
        L[k + 1:, k] = l
        M[k + 1:, k + 1:] -= np.outer(l, M[k + 1:, k])
    return L, d, nfloor


def solve(L, d, x):
    m = len(x)
    y = x.copy()
    for k in range(m):
        y[k + 1:] -= L[k + 1:, k] * y[k]
    y /= d
    for k in range(m - 1, -1, -1):
        y[k] -= L[k + 1:, k] @ y[k + 1:]
    return y


def noda_pixel(H, states, warm, policy):
    """Returns nbar, stats.  warm: dict key->value of the previous pixel's vectors (all sectors)."""
    kappa, theta_tol, warm_fill = policy
    tc = states.sum(axis=1)
    keys = [tuple(s) for s in states]
    scale = np.abs(H).sum(axis=1).max() + 1e-300
    secs = {}
    for c in np.unique(tc):
        idx = np.where(tc == c)[0]
        A = -np.abs(H[np.ix_(idx, idx)])
        A[np.arange(len(idx)), np.arange(len(idx))] = np.diag(H)[idx]
        x = np.array([warm.get(keys[i], 0.0) for i in idx])
        if (x > 0).any():
            x = np.maximum(x, warm_fill * x.max())
        else:
            x = np.ones(len(idx))
        x = x / np.linalg.norm(x)
        w = A @ x
        gersh = (2 * np.diag(A) - np.abs(A).sum(axis=1)).min()
        secs[c] = dict(idx=idx, A=A, x=x, lo=max(gersh, (w / x).min()), ub=x @ w, live=True, done=False, L=None, nf=0, sig=None)
    stats = dict(fac=0, sol=0, steps=0, maxlen=0)
    for it in range(10):
        U = min(s["ub"] for s in secs.values())
        any_work = False
        step_len = 0
        for c, s in secs.items():
            if s["live"] and s["lo"] > U + 1e-13 * scale:
                s["live"] = False
            if not s["live"] or s["done"]:
                continue
            any_work = True
        if not any_work:
            break
        did_fac = False
        for c, s in secs.items():
            if not s["live"] or s["done"]:
                continue
            m = len(s["idx"])
            if m == 1:
                s["lo"] = s["ub"] = s["A"][0, 0]
                s["done"] = True
                continue
            step_len = max(step_len, m)
            if s["L"] is None or (s["ub"] - s["lo"]) < kappa * (s["ub"] - s["sig"]):
                s["L"], s["d"], nfl = ldl_solve(s["A"], s["lo"], s["x"], 1e-15 * scale)
                s["sig"] = s["lo"]
                s["nf"] += 1
                did_fac = True
                if nfl:
                    s["done"] = True
            y = solve(s["L"], s["d"], s["x"])
            r = s["x"] / y
            if not s["done"]:
                s["lo"] = max(s["lo"], s["sig"] + r.min())
            ny2 = y @ y
            xy = y @ s["x"]
            s["ub"] = min(s["ub"], s["sig"] + xy / ny2)
            sin2 = max(0.0, 1.0 - xy * xy / ny2)
            s["x"] = y / np.sqrt(ny2)
            delta = xy / ny2
            if sin2 < theta_tol ** 2 and delta < 1e-3 * scale:
                s["done"] = True
        stats["steps"] += 1
        stats["fac"] += did_fac
        stats["sol"] += 1
        stats["maxlen"] = max(stats["maxlen"], step_len)
    best = min(secs.values(), key=lambda s: s["ub"])
    psi2 = best["x"] ** 2
    nbar = psi2 @ states[best["idx"]].astype(float)
    new_warm = {}
    for s in secs.values():
        for i, v in zip(s["idx"], s["x"]):
            new_warm[keys[i]] = v
    return nbar, stats, new_warm


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-dot", type=int, default=4)
    ap.add_argument("--envs", type=int, default=4)
    ap.add_argument("--rows", type=int, default=4)
    ap.add_argument("--kappa", type=float, default=0.25)
    ap.add_argument("--fill", type=float, default=0.3)
    ap.add_argument("--tol", type=float, default=3e-8)
    args = ap.parse_args()
    import bench
    from util import oracle_model, oracle_scan
    dev, mb, sets = bench.build_workload(args.envs, args.n_dot, 64, 0, 1, "B")
    scans = sets[0]
    tot = dict(fac=0, sol=0, steps=0, pix=0)
    worst = 0.0
    nbad = 0
    lens = []
    for si in range(len(scans)):
        rec = scans[si]
        m = oracle_model(mb, int(rec["env_id"]), 0)
        s0 = oracle_scan(rec, mb.n_volt, 0)
        grid = composer.affine_grid(s0.v0, s0.dx, s0.dy, s0.nx, s0.ny).reshape(s0.ny, s0.nx, mb.n_volt)
        for iy in range(0, s0.ny, max(1, s0.ny // args.rows)):
            v = grid[iy]
            cinv = np.asarray(m.cdd_inv, float)
            g = v @ np.asarray(m.cgd, float).T
            n_c = path_b.continuous_ground_state(g, cinv, None)
            st = path_b.select_charge_states(g, n_c, cinv, m.num_charge_states, m.charge_state_batch_size)
            t = path_b.tunnel_couplings(m, v)
            h, f = path_b.hamiltonian(st, g, cinv, t)
            w, vec = np.linalg.eigh(h)
            ref = np.einsum("pm,pmd->pd", vec[:, :, 0] ** 2, st.astype(float))
            gap = w[:, 1] - w[:, 0]
            warm = {}
            for p in range(len(v)):
                nbar, stt, warm = noda_pixel(h[p], st[p], warm, (args.kappa, args.tol, args.fill))
                err = np.abs(nbar - ref[p]).max()
                if gap[p] > 1e-5:
                    worst = max(worst, err)
                    if err > 1e-7:
                        nbad += 1
                        print("bad", si, iy, p, err, gap[p], stt)
                for k in ("fac", "sol", "steps"):
                    tot[k] += stt[k]
                tot["pix"] += 1
                lens.append(stt["maxlen"])
    print({k: v / tot["pix"] for k, v in tot.items() if k != "pix"}, "pixels", tot["pix"], "worst err", worst, "bad", nbad,
          "mean maxlen", np.mean(lens))


if __name__ == "__main__":
    main()
