import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import numpy as np
import ipm_exp as E

def ipm_w(H, q, G, h, ws=None, o=None, tol=1e-11, max_iter=50):
    o = o or {}
    n, m = q.size, h.size; nz = n - 1
    x = np.zeros(n)
    L = np.linalg.cholesky(H[:nz, :nz]); x[:nz] = -np.linalg.solve(L.T, np.linalg.solve(L, q[:nz]))
    slack0 = h - G @ x
    hscale = 1.0 + np.abs(h).max()
    if slack0.min() >= -1e-12 * hscale: return x, 0, None, 0
    qs = 1.0 + np.abs(q).max()
    mu0 = max(1e-2 * qs * hscale / m, 1e-8)
    s = np.maximum(slack0, 1e-2 * hscale); lam = mu0 / s
    mode = o.get("mode", "cur")
    if ws is not None and mode != "cold":
        xw, lw = ws
        if mode == "cur":
            x = xw.copy(); sl = h - G @ x
            s = np.maximum(sl, 1e-2 * hscale); lam = np.maximum(lw, 1e-4 * qs / hscale)
        elif mode == "xonly":
            x = xw.copy(); sl = h - G @ x
            s = np.maximum(sl, o.get("sf", 1e-2) * hscale); lam = o.get("mf", 1.0) * mu0 / s
        elif mode == "clip":
            x = xw.copy(); sl = h - G @ x
            s = np.maximum(sl, o.get("sf", 1e-2) * hscale)
            lam = np.maximum(lw, 1e-4 * qs / hscale)
            mut = o.get("mt", 1.0) * (s @ lam) / m
            lo, hi = o.get("lo", 0.1), o.get("hi", 10.0)
            lam = np.clip(lam, lo * mut / s, hi * mut / s)
        elif mode == "blend":
            # convex combination of the warm point and the cold point in x; centred multipliers
            th = o.get("th", 0.5)
            x = th * xw + (1 - th) * x; sl = h - G @ x
            s = np.maximum(sl, 1e-2 * hscale); lam = np.maximum(th * lw, mu0 / s)
    best = 1e300; tol_mu = 1e-3 * tol
    for it in range(max_iter + 1):
        Hxq = H @ x + q; Gl = G.T @ lam
        rd = Hxq + Gl; rp = G @ x + s - h; mu = s @ lam / m
        qd = qs + max(np.abs(Hxq).max(), np.abs(Gl).max())
        merit = max(np.abs(rd).max() / (tol * qd), np.abs(rp).max() / (tol * hscale), mu * m / (tol_mu * qs * hscale))
        if merit <= 1.0 or (best <= 1e3 and merit >= best): return x, it, lam, 0
        best = min(best, merit)
        if it == max_iter: return x, it, lam, 1
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
    return x, it, lam, 1

def run(o, label):
    its = np.zeros((E.T, E.N)); err = 0; bad = 0
    wsx = [None] * E.N
    for rec in E.data:
        i, k = rec["i"], rec["k"]
        G, h = rec["G"][:-1], rec["h"][:-1]
        x, it, lam, st = ipm_w(rec["H"], rec["q"], G, h, ws=wsx[i], o=o)
        its[k, i] = it; bad += st
        err = max(err, np.abs(x - rec["x"]).max() / (1 + np.abs(rec["x"]).max()))
        wsx[i] = (x, lam) if (lam is not None and st == 0) else None
    a = its[5:]
    print("%-40s its mean %.2f p90 %d p99 %d max %d permax %.1f | err %.1e bad %d" % (label, a.mean(), np.percentile(a, 90), np.percentile(a, 99), a.max(), a.max(axis=1).mean(), err, bad), flush=True)

if __name__ == "__main__":
    run(dict(mode="cur"), "current")
    run(dict(mode="cold"), "cold")
    run(dict(mode="xonly"), "x only, centred lam")
    run(dict(mode="xonly", mf=0.1), "x only, mu0*0.1")
    run(dict(mode="xonly", mf=0.01, sf=1e-3), "x only, mu0*0.01 sf 1e-3")
    run(dict(mode="clip"), "clip .1..10")
    run(dict(mode="clip", lo=0.3, hi=3), "clip .3..3")
    run(dict(mode="clip", lo=0.1, hi=10, mt=0.1), "clip .1..10 mt .1")
    run(dict(mode="blend", th=0.5), "blend .5")
    run(dict(mode="blend", th=0.9), "blend .9")
