"""Active-set-first study: before the interior-point iteration, try the previous period's active set as
equalities (method of multipliers on H + rho G_A' G_A: ONE factorisation, a few solves), verify the KKT
conditions of the full QP, and fall back to the warm-started IPM when the verification fails."""
import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import pickle
import numpy as np
import ipm_exp as E
from ipm_exp8 import ipm_w

def active_set_try(H, q, G, h, xw, lw, qs, hscale, rho_rel=1e7, nit=3, tol=1e-11, thr=1e-6, second=False):
    m = h.size
    act = lw > thr * qs / hscale
    cost = 0.0
    for attempt in range(2 if second else 1):
        if not act.any():
            return None, None, cost
        GA, hA = G[act], h[act]
        rho = rho_rel * np.abs(np.diag(H)).max() / max((GA * GA).sum(axis=1).max(), 1e-300)
        Phi = H + rho * GA.T @ GA
        Lc = E.gchol(Phi)
        solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        lam = np.where(act, lw, 0.0)[act] if attempt == 0 else np.maximum(lamfull[act], 0.0)
        cost += 1.0
        for _ in range(nit):
            x = solve(-q - GA.T @ lam + rho * GA.T @ hA)
            lam = lam + rho * (GA @ x - hA)
            cost += 0.25
        lamfull = np.zeros(m); lamfull[act] = lam
        Hxq = H @ x + q; Gl = G.T @ lamfull
        qd = qs + max(np.abs(Hxq).max(), np.abs(Gl).max())
        rd = np.abs(Hxq + Gl).max()
        viol = G @ x - h
        ok = (rd <= tol * qd and viol.max() <= tol * hscale and np.abs(viol[act]).max() <= tol * hscale
              and lam.min() >= -tol * qs)
        if ok:
            return x, lamfull, cost
        # second attempt: add violated rows, drop rows with negative multipliers
        newact = (act & (lamfull > 0)) | (viol > tol * hscale)
        if (newact == act).all():
            break
        act = newact
    return None, None, cost

def run(label, **kw):
    its = np.zeros((E.T, E.N)); cost = np.zeros((E.T, E.N)); err = 0; bad = 0; ok = 0; tried = 0
    wsx = [None] * E.N
    for rec in E.data:
        i, k = rec["i"], rec["k"]
        G, h = rec["G"][:-1], rec["h"][:-1]
        H, q = rec["H"], rec["q"]
        n = q.size; nz = n - 1
        x0 = np.zeros(n); L = np.linalg.cholesky(H[:nz, :nz]); x0[:nz] = -np.linalg.solve(L.T, np.linalg.solve(L, q[:nz]))
        hscale = 1.0 + np.abs(h).max(); qs = 1.0 + np.abs(q).max()
        c = 0.0; x = None
        if (h - G @ x0).min() >= -1e-12 * hscale:
            x, lam, it = x0, None, 0
        else:
            if wsx[i] is not None and kw.get("on", True):
                tried += 1
                x, lam, c = active_set_try(H, q, G, h, wsx[i][0], wsx[i][1], qs, hscale, **{k_: v for k_, v in kw.items() if k_ != "on"})
                if x is not None:
                    ok += 1; it = 0
            if x is None:
                x, it, lam, st = ipm_w(H, q, G, h, ws=wsx[i], o=dict(mode="cur"))
                bad += st
        its[k, i] = it; cost[k, i] = it + c
        err = max(err, np.abs(x - rec["x"]).max() / (1 + np.abs(rec["x"]).max()))
        wsx[i] = (x, lam) if lam is not None else None
    a = cost[5:]
    print("%-34s cost mean %.2f p90 %.1f p99 %.1f max %.1f permax %.1f | AS ok %d/%d | err %.1e bad %d" % (
        label, a.mean(), np.percentile(a, 90), np.percentile(a, 99), a.max(), a.max(axis=1).mean(), ok, tried, err, bad), flush=True)
    pm = a.max(axis=1)
    print("    per-period max cost:", " ".join("%.0f" % v for v in pm))
    return cost

if __name__ == "__main__":
    E.data = pickle.load(open(sys.argv[1], "rb")); E.N = int(sys.argv[2]); E.T = int(sys.argv[3])
    run("IPM only (current)", on=False)
    run("AS first, 1 attempt")
    run("AS first, 2 attempts", second=True)
    run("AS first, rho 1e5", rho_rel=1e5)
    run("AS first, rho 1e9", rho_rel=1e9)
