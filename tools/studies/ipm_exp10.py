"""Crossover study: inside the IPM, from iteration K0 on and every STEP iterations, guess the active set from the
iterate (lam_r / lscale > s_r / hscale), solve the equality-constrained QP on it (method of multipliers, ONE
factorisation of H + rho G_A'G_A), verify the KKT conditions of the full QP and stop when they hold."""
import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import pickle
import numpy as np
import ipm_exp as E

def crossover(H, q, G, h, x, s, lam, qs, hscale, tol, rho_rel=1e6, nit=3, kappa=1.0):
    act = lam * hscale > kappa * s * qs
    if not act.any(): return None
    GA, hA = G[act], h[act]
    rho = rho_rel * np.median(np.abs(np.diag(H))) / max((GA * GA).sum(axis=1).max(), 1e-300)
    Lc = E.gchol(H + rho * GA.T @ GA)
    solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
    l = lam[act].copy()
    for _ in range(nit):
        xn = solve(-q - GA.T @ l + rho * GA.T @ hA)
        l = l + rho * (GA @ xn - hA)
    lf = np.zeros(h.size); lf[act] = l
    Hxq = H @ xn + q; Gl = G.T @ lf
    qd = qs + max(np.abs(Hxq).max(), np.abs(Gl).max())
    viol = G @ xn - h
    ok = (np.abs(Hxq + Gl).max() <= tol * qd and viol.max() <= tol * hscale and np.abs(viol[act]).max() <= tol * hscale
          and l.min() >= -tol * qs)
    return (xn, lf) if ok else None

def ipm_x(H, q, G, h, ws=None, tol=1e-11, max_iter=50, K0=5, STEP=2, cross=True, ccost=1.5, **ck):
    n, m = q.size, h.size; nz = n - 1
    x = np.zeros(n)
    L = np.linalg.cholesky(H[:nz, :nz]); x[:nz] = -np.linalg.solve(L.T, np.linalg.solve(L, q[:nz]))
    slack0 = h - G @ x
    hscale = 1.0 + np.abs(h).max()
    if slack0.min() >= -1e-12 * hscale: return x, 0, None, 0, 0.0
    qs = 1.0 + np.abs(q).max()
    mu0 = max(1e-2 * qs * hscale / m, 1e-8)
    s = np.maximum(slack0, 1e-2 * hscale); lam = mu0 / s
    if ws is not None:
        xw, lw = ws
        x = xw.copy(); sl = h - G @ x
        s = np.maximum(sl, 1e-2 * hscale); lam = np.maximum(lw, 1e-4 * qs / hscale)
    best = 1e300; tol_mu = 1e-3 * tol; extra = 0.0
    for it in range(max_iter + 1):
        Hxq = H @ x + q; Gl = G.T @ lam
        rd = Hxq + Gl; rp = G @ x + s - h; mu = s @ lam / m
        qd = qs + max(np.abs(Hxq).max(), np.abs(Gl).max())
        merit = max(np.abs(rd).max() / (tol * qd), np.abs(rp).max() / (tol * hscale), mu * m / (tol_mu * qs * hscale))
        if merit <= 1.0 or (best <= 1e3 and merit >= best): return x, it, lam, 0, extra
        best = min(best, merit)
        if it == max_iter: return x, it, lam, 1, extra
        if cross and it >= K0 and (it - K0) % STEP == 0:
            extra += ccost
            r = crossover(H, q, G, h, x, s, lam, qs, hscale, tol, **ck)
            if r is not None: return r[0], it, r[1], 0, extra
        d = lam / s
        Lc = E.gchol(H + G.T @ (d[:, None] * G)); solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        dx = solve(-rd - G.T @ (d * rp - lam)); ds = -rp - G @ dx; dl = -lam - d * ds
        aa = E.alpha_max(s, ds, lam, dl)
        mu_a = (s + aa * ds) @ (lam + aa * dl) / m
        ratio = mu_a / mu; sig = ratio ** 3
        rc = s * lam + ds * dl - sig * mu
        dx = solve(-rd - G.T @ ((lam * rp - rc) / s)); ds = -rp - G @ dx; dl = -(rc + lam * ds) / s
        am = E.alpha_max(s, ds, lam, dl)
        tau = min(max(0.99, 1.0 - ratio), 1 - 1e-6)
        a = min(1.0, tau * am)
        x, s, lam = x + a * dx, s + a * ds, lam + a * dl
    return x, it, lam, 1, extra

def run(label, **kw):
    cost = np.zeros((E.T, E.N)); bad = 0; err = 0.0
    wsx = [None] * E.N; ref = E.ref
    for j, rec in enumerate(E.data):
        i, k = rec["i"], rec["k"]
        G, h = rec["G"][:-1], rec["h"][:-1]
        x, it, lam, st, extra = ipm_x(rec["H"], rec["q"], G, h, ws=wsx[i], **kw)
        cost[k, i] = it + extra; bad += st
        if ref is not None: err = max(err, np.abs(x - ref[j]).max() / (1 + np.abs(ref[j]).max()))
        wsx[i] = (x, lam) if (lam is not None and st == 0) else None
    a = cost[5:]
    print("%-36s cost mean %.2f p90 %.1f p99 %.1f max %.1f permax %.1f | err vs baseline %.1e bad %d" % (
        label, a.mean(), np.percentile(a, 90), np.percentile(a, 99), a.max(), a.max(axis=1).mean(), err, bad), flush=True)

if __name__ == "__main__":
    E.data = pickle.load(open(sys.argv[1], "rb")); E.N = int(sys.argv[2]); E.T = int(sys.argv[3]); E.ref = None
    # reference solutions: the current IPM at its own tolerance
    wsx = [None] * E.N; ref = []
    for rec in E.data:
        x, it, lam, st, _ = ipm_x(rec["H"], rec["q"], rec["G"][:-1], rec["h"][:-1], ws=wsx[rec["i"]], cross=False)
        ref.append(x); wsx[rec["i"]] = (x, lam) if (lam is not None and st == 0) else None
    E.ref = ref
    run("IPM only (current)", cross=False)
    run("crossover K0=5 step 2 rho 1e4", rho_rel=1e4)
    run("crossover K0=5 step 2 rho 1e4 nit 5", rho_rel=1e4, nit=5)
    run("crossover K0=5 step 2 rho 1e6", rho_rel=1e6)
    run("crossover K0=7 step 3 rho 1e4 nit 5", K0=7, STEP=3, rho_rel=1e4, nit=5)
    run("crossover K0=4 step 2 rho 1e3 nit 5", K0=4, STEP=2, rho_rel=1e3, nit=5)
