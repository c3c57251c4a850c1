import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import numpy as np
import ipm_exp as E

def ipm2(H, q, G, h, ws=None, o=None, tol=1e-11, max_iter=50, trace=False):
    n, m = q.size, h.size; nz = n - 1
    x = np.zeros(n)
    L = np.linalg.cholesky(H[:nz, :nz]); x[:nz] = -np.linalg.solve(L.T, np.linalg.solve(L, q[:nz]))
    slack0 = h - G @ x
    hscale = 1.0 + np.abs(h).max()
    if slack0.min() >= -1e-12 * hscale: return x, 0, None, 0
    qs = 1.0 + np.abs(q).max()
    viol = max(0.0, -slack0.min())
    mode = o.get("init", "base")
    if mode == "base":
        mu0 = max(1e-2 * qs * hscale / m, 1e-8)
        s = np.maximum(slack0, 1e-2 * hscale); lam = mu0 / s
    elif mode == "shift":
        s = np.maximum(slack0, 0) + o.get("sh", 1.0) * viol + 1e-2 * hscale
        mu0 = o.get("mu0", 1e-2) * qs * hscale / m
        lam = mu0 / s
    elif mode == "x0zero":
        x = np.zeros(n); slack0 = h - G @ x
        s = np.maximum(slack0, o.get("sf", 1e-2) * hscale)
        mu0 = o.get("mu0", 1e-2) * qs * hscale / m
        lam = mu0 / s
    elif mode == "mehrotra":
        # one affine step from (x0, s=max(slack,.), lam=1-ish) then shift
        s = np.maximum(slack0, 1e-2 * hscale); lam = np.full(m, max(1e-2 * qs / hscale, 1e-8))
        d = lam / s; rd = H @ x + q + G.T @ lam; rp = G @ x + s - h
        Lc = E.gchol(H + G.T @ (d[:, None] * G)); solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        dx = solve(-rd - G.T @ (d * rp - lam)); ds = -rp - G @ dx; dl = -lam - d * ds
        x = x + dx; s2 = s + ds; l2 = lam + dl
        s = np.maximum(np.abs(s2), 1e-2 * hscale); lam = np.maximum(np.abs(l2), 1e-8)
    if ws is not None:
        xw, lw = ws
        x = xw.copy(); sl = h - G @ x
        s = np.maximum(sl, o.get("wsf", 1e-2) * hscale); lam = np.maximum(lw, o.get("wlmin", 1e-4) * qs / hscale)
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
        sig = (mu_a / mu) ** 3
        rc = s * lam + ds * dl - sig * mu
        dx = solve(-rd - G.T @ ((lam * rp - rc) / s)); ds = -rp - G @ dx; dl = -(rc + lam * ds) / s
        am = E.alpha_max(s, ds, lam, dl)
        tau = max(0.99, 1.0 - mu_a / mu) if o.get("adapt", True) else 0.99
        a = min(1.0, tau * am)
        if trace: print("%2d rd %.1e rp %.1e mu %.1e | a_aff %.3f sig %.1e a %.4f" % (it, np.abs(rd).max(), np.abs(rp).max(), mu, aa, sig, a))
        x, s, lam = x + a * dx, s + a * ds, lam + a * dl
    return x, it, lam, 1

def run(o, label, cold=False):
    its = np.zeros((E.T, E.N), int); err = 0; bad = 0
    wsx = [None] * E.N
    for rec in E.data:
        i, k = rec["i"], rec["k"]
        G, h = rec["G"][:-1], rec["h"][:-1]
        x, it, lam, st = ipm2(rec["H"], rec["q"], G, h, ws=None if cold else wsx[i], o=o)
        its[k, i] = it; bad += st
        err = max(err, np.abs(x - rec["x"]).max() / (1 + np.abs(rec["x"]).max()))
        wsx[i] = (x, lam) if (lam is not None and st == 0) else None
    a = its[5:]
    print("%-40s mean %.2f p90 %d p99 %d max %d  permax-mean %.1f err %.1e bad %d" % (label, a.mean(), np.percentile(a, 90), np.percentile(a, 99), a.max(), a.max(axis=1).mean(), err, bad), flush=True)
    return its
if __name__ == "__main__":
    run(dict(init="base"), "cold base", cold=True)
    for sh in (0.5, 1.0, 2.0):
        for mu0 in (1e-2, 1e-1, 1.0):
            run(dict(init="shift", sh=sh, mu0=mu0), "cold shift sh=%g mu0=%g" % (sh, mu0), cold=True)
    run(dict(init="x0zero"), "cold x0zero", cold=True)
    run(dict(init="x0zero", mu0=1e-1), "cold x0zero mu0=.1", cold=True)
    run(dict(init="mehrotra"), "cold mehrotra", cold=True)
