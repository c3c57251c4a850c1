import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import numpy as np
import ipm_exp as E

def ipm3(H, q, G, h, ws=None, o=None, max_iter=50, trace=False):
    tol = o.get("tol", 1e-11); tol_mu = o.get("tmf", 1e-3) * tol
    n, m = q.size, h.size; nz = n - 1
    x = np.zeros(n)
    L = np.linalg.cholesky(H[:nz, :nz]); x[:nz] = -np.linalg.solve(L.T, np.linalg.solve(L, q[:nz]))
    slack0 = h - G @ x
    hscale = 1.0 + np.abs(h).max()
    if slack0.min() >= -1e-12 * hscale: return x, 0, None, 0, 0
    qs = 1.0 + np.abs(q).max()
    extra = 0
    def meh(x, s, lam):
        d = lam / s; rd = H @ x + q + G.T @ lam; rp = G @ x + s - h
        Lc = E.gchol(H + G.T @ (d[:, None] * G)); solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        dx = solve(-rd - G.T @ (d * rp - lam)); ds = -rp - G @ dx; dl = -lam - d * ds
        x = x + dx; s2 = s + ds; l2 = lam + dl
        return x, np.maximum(np.abs(s2), o.get("sfl", 1e-2) * hscale), np.maximum(np.abs(l2), o.get("lfl", 1e-8))
    if ws is None or o.get("nowarm"):
        s = np.maximum(slack0, 1e-2 * hscale); lam = np.full(m, max(1e-2 * qs / hscale, 1e-8))
        x, s, lam = meh(x, s, lam); extra = 1
    else:
        xw, lw = ws
        x = xw.copy(); sl = h - G @ x
        s = np.maximum(sl, o.get("wsf", 1e-2) * hscale); lam = np.maximum(lw, o.get("wlmin", 1e-4) * qs / hscale)
        if o.get("wmeh"):
            x, s, lam = meh(x, s, lam); extra = 1
    best = 1e300
    for it in range(max_iter + 1):
        Hxq = H @ x + q; Gl = G.T @ lam
        rd = Hxq + Gl; rp = G @ x + s - h; mu = s @ lam / m
        qd = qs + max(np.abs(Hxq).max(), np.abs(Gl).max())
        merit = max(np.abs(rd).max() / (tol * qd), np.abs(rp).max() / (tol * hscale), mu * m / (tol_mu * qs * hscale))
        if merit <= 1.0 or (best <= 1e3 and merit >= best): return x, it, lam, 0, extra
        best = min(best, merit)
        if it == max_iter: return x, it, lam, 1, extra
        d = lam / s
        Lc = E.gchol(H + G.T @ (d[:, None] * G)); solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        dx = solve(-rd - G.T @ (d * rp - lam)); ds = -rp - G @ dx; dl = -lam - d * ds
        aa = E.alpha_max(s, ds, lam, dl)
        mu_a = (s + aa * ds) @ (lam + aa * dl) / m
        sig = (mu_a / mu) ** 3
        rc = s * lam + ds * dl - sig * mu
        dx = solve(-rd - G.T @ ((lam * rp - rc) / s)); ds = -rp - G @ dx; dl = -(rc + lam * ds) / s
        am = E.alpha_max(s, ds, lam, dl)
        tau = max(0.99, 1.0 - mu_a / mu)
        a = min(1.0, tau * am)
        if trace: print("%2d rd %.1e rp %.1e mu %.1e | a_aff %.3f sig %.1e a %.4f" % (it, np.abs(rd).max(), np.abs(rp).max(), mu, aa, sig, a))
        x, s, lam = x + a * dx, s + a * ds, lam + a * dl
    return x, it, lam, 1, extra

def run(o, label):
    its = np.zeros((E.T, E.N), float); errs = []; bad = 0; jerr = 0
    wsx = [None] * E.N
    for rec in E.data:
        i, k = rec["i"], rec["k"]
        G, h = rec["G"][:-1], rec["h"][:-1]
        x, it, lam, st, extra = ipm3(rec["H"], rec["q"], G, h, ws=wsx[i], o=o)
        its[k, i] = it + 0.7 * extra; bad += st
        errs.append(np.abs(x - rec["x"]).max() / (1 + np.abs(rec["x"]).max()))
        J = 0.5 * x @ rec["H"] @ x + rec["q"] @ x; Js = 0.5 * rec["x"] @ rec["H"] @ rec["x"] + rec["q"] @ rec["x"]
        jerr = max(jerr, abs(J - Js) / (1 + abs(Js)))
        wsx[i] = (x, lam) if (lam is not None and st == 0) else None
    a = its[5:]; errs = np.array(errs)
    print("%-44s mean %.2f p90 %.1f p99 %.1f max %.1f permax-mean %.1f | err max %.1e p99 %.1e Jerr %.1e bad %d" % (label, a.mean(), np.percentile(a, 90), np.percentile(a, 99), a.max(), a.max(axis=1).mean(), errs.max(), np.percentile(errs, 99), jerr, bad), flush=True)
    return its
if __name__ == "__main__":
    for tol in (1e-11, 1e-10, 1e-9, 1e-8):
        for tmf in (1e-3, 1.0):
            run(dict(tol=tol, tmf=tmf, nowarm=True), "cold tol=%g tmf=%g" % (tol, tmf))
    for tol in (1e-11, 1e-9):
        run(dict(tol=tol, tmf=1.0), "warm tol=%g tmf=1" % tol)
        run(dict(tol=tol, tmf=1.0, wmeh=True), "warm+meh tol=%g tmf=1" % tol)
