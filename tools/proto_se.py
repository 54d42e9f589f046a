"""NumPy prototype of a breadth-first lattice enumeration for the select stage (design study, not shipped): node counts per
level when all candidates with E <= tau (tau = 32nd best energy, i.e. a perfect warm start) are enumerated level by level
on C = L D L^T."""
import sys
import numpy as np

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests")
sys.path.insert(0, "/root/repo/rl-agent-for-qubit-array-tuning_b200")
from oracle import path_b, composer  # noqa: E402


def ldl(c):
    n = len(c)
    a = c.copy()
    L = np.eye(n)
    d = np.zeros(n)
    for k in range(n):
        d[k] = a[k, k]
        L[k + 1:, k] = a[k + 1:, k] / d[k]
        a[k + 1:, k + 1:] -= np.outer(L[k + 1:, k], a[k + 1:, k])
    return L, d


def enumerate_levels(cinv, r, f, tau, order, use_rem=True):
    """order: permutation of dots, order[-1] is decided first.  Returns node counts per level and the leaves."""
    n = len(r)
    P = np.eye(n)[order]
    c = P @ cinv @ P.T
    rp, fp = r[order], f[order]
    L, d = ldl(c)
    rc = L.T @ rp
    rem = np.zeros(n)
    if use_rem:
        lb = np.zeros(n)
        for k in range(n):
            lo = np.where(fp[k + 1:] <= 0, 0.0, -1.0)
            l = L[k + 1:, k]
            smin = np.minimum(l * lo, l * 2).sum()
            smax = np.maximum(l * lo, l * 2).sum()
            ylo = rc[k] + (0.0 if fp[k] <= 0 else -1.0) + smin
            yhi = rc[k] + 2.0 + smax
            dist = ylo if ylo > 0 else (-yhi if yhi < 0 else 0.0)
            lb[k] = d[k] * dist * dist
        rem = np.concatenate([[0.0], np.cumsum(lb)[:-1]])
    nodes = [((), 0.0)]
    counts = []
    for lev in range(n):
        k = n - 1 - lev
        new = []
        for digs, pe in nodes:
            # digs: deltas of dots k+1..n-1 in order (k+1 first)
            z_hi = np.array([rp[k + 1 + i] + digs[i] for i in range(len(digs))])
            ck = rp[k] + (L[k + 1:, k] @ z_hi if len(digs) else 0.0)
            for dg in (-1, 0, 1, 2):
                if fp[k] + dg < 0:
                    continue
                y = ck + dg
                pe2 = pe + d[k] * y * y
                if pe2 + rem[k] * (1 - 1e-9) <= tau * (1 + 1e-12) + 1e-300:
                    new.append(((dg,) + digs, pe2))
        nodes = new
        counts.append(len(nodes))
    return counts, nodes


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-dot", type=int, default=8)
    ap.add_argument("--envs", type=int, default=2)
    ap.add_argument("--rows", type=int, default=2)
    ap.add_argument("--stale", type=int, default=1, help="tau taken from the basis of the pixel this many steps back")
    args = ap.parse_args()
    import bench
    from util import oracle_model, oracle_scan
    dev, mb, sets = bench.build_workload(args.envs, args.n_dot, 64, 0, 1, "B")
    scans = sets[0]
    allc = {"nat": [], "emptytop": [], "rev": []}
    leaves = []
    for si in range(len(scans)):
        rec = scans[si]
        m = oracle_model(mb, int(rec["env_id"]), 0)
        s0 = oracle_scan(rec, mb.n_volt, 0)
        grid = composer.affine_grid(s0.v0, s0.dx, s0.dy, s0.nx, s0.ny).reshape(s0.ny, s0.nx, mb.n_volt)
        cinv = np.asarray(m.cdd_inv, float)
        for iy in range(0, s0.ny, max(1, s0.ny // args.rows)):
            v = grid[iy]
            g = v @ np.asarray(m.cgd, float).T
            n_c = path_b.continuous_ground_state(g, cinv, None)
            st = path_b.select_charge_states(g, n_c, cinv, 32, m.charge_state_batch_size)
            fl = np.floor(n_c)
            for p in range(args.stale, len(v)):
                # warm start: the previous pixel's states re-evaluated here (those still inside the candidate box)
                prev = st[p - args.stale].astype(float)
                ok = ((prev - fl[p] >= -1) & (prev - fl[p] <= 2)).all(axis=1)
                zz = prev[ok] - g[p]
                e = np.einsum("mi,ij,mj->m", zz, cinv, zz)
                if ok.sum() < 32:
                    tau = np.inf
                else:
                    tau = np.sort(e)[31]
                if not np.isfinite(tau):
                    continue
                r = fl[p] - g[p]
                n = len(r)
                empty = fl[p] <= 0
                nat = np.arange(n)
                orders = {"nat": nat, "emptytop": np.concatenate([nat[~empty], nat[empty][np.argsort(-g[p][empty])]]),
                          "rev": nat[::-1]}
                for name, od in orders.items():
                    cnt, nodes = enumerate_levels(cinv, r, fl[p], tau, od)
                    allc[name].append(cnt)
                    if name == "nat":
                        leaves.append(len(nodes))
    for name, c in allc.items():
        c = np.array(c)
        print(name, "mean per level", np.round(c.mean(axis=0), 1), "max per level", c.max(axis=0), "sum mean", c.sum(axis=1).mean(),
              "p99 sum", np.quantile(c.sum(axis=1), 0.99))
    print("leaves mean", np.mean(leaves), "max", np.max(leaves), "pixels", len(leaves))


if __name__ == "__main__":
    main()
