import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import numpy as np
import ipm_exp as E

def ipm_g(H, q, G, h, ws=None, o=None, tol=1e-11, max_iter=50):
    o = o or {}
    K = o.get("K", 0); delta = o.get("delta", 0.3); gam = o.get("gamma", 0.1); bmin = o.get("bmin", 0.1); bmax = o.get("bmax", 10.0)
    n, m = q.size, h.size; nz = n - 1
    x = np.zeros(n)
    L = np.linalg.cholesky(H[:nz, :nz]); x[:nz] = -np.linalg.solve(L.T, np.linalg.solve(L, q[:nz]))
    slack0 = h - G @ x
    hscale = 1.0 + np.abs(h).max()
    if slack0.min() >= -1e-12 * hscale: return x, 0, None, 0, 0
    qs = 1.0 + np.abs(q).max()
    mu0 = max(1e-2 * qs * hscale / m, 1e-8)
    viol = max(0.0, -slack0.min())
    if o.get("cold") == "viol":
        s = np.maximum(slack0, max(1e-2 * hscale, o.get("vf", 1.0) * viol)); lam = mu0 / s
    else:
        s = np.maximum(slack0, 1e-2 * hscale); lam = mu0 / s
    if ws is not None:
        xw, lw = ws
        x = xw.copy(); sl = h - G @ x
        s = np.maximum(sl, 1e-2 * hscale); lam = np.maximum(lw, 1e-4 * qs / hscale)
    best = 1e300; tol_mu = 1e-3 * tol; ncorr = 0
    for it in range(max_iter + 1):
        Hxq = H @ x + q; Gl = G.T @ lam
        rd = Hxq + Gl; rp = G @ x + s - h; mu = s @ lam / m
        qd = qs + max(np.abs(Hxq).max(), np.abs(Gl).max())
        merit = max(np.abs(rd).max() / (tol * qd), np.abs(rp).max() / (tol * hscale), mu * m / (tol_mu * qs * hscale))
        if merit <= 1.0 or (best <= 1e3 and merit >= best): return x, it, lam, 0, ncorr
        best = min(best, merit)
        if it == max_iter: return x, it, lam, 1, ncorr
        d = lam / s
        Lc = E.gchol(H + G.T @ (d[:, None] * G)); solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        dx = solve(-rd - G.T @ (d * rp - lam)); ds = -rp - G @ dx; dl = -lam - d * ds
        aa = E.alpha_max(s, ds, lam, dl)
        mu_a = (s + aa * ds) @ (lam + aa * dl) / m
        ratio = mu_a / mu; sig = ratio ** 3
        rc = s * lam + ds * dl - sig * mu
        dx = solve(-rd - G.T @ ((lam * rp - rc) / s)); ds = -rp - G @ dx; dl = -(rc + lam * ds) / s
        am = E.alpha_max(s, ds, lam, dl)
        for kc in range(K):
            if am >= o.get("skip_above", 0.9): break
            at = min(1.0, am + delta) if o.get("additive", True) else min(1.0, am * o.get("mult", 2.0))
            v = (s + at * ds) * (lam + at * dl)
            mut = sig * mu if o.get("target", "sigmu") == "sigmu" else max(sig * mu, 0.0)
            mut = max(mut, o.get("mutfloor", 0.0) * mu)
            t = np.clip(v, bmin * mut, bmax * mut) - v
            t = np.maximum(t, -bmax * mut)
            # correction: rd=0, rp=0, s*dl + lam*ds = t  ->  dx = solve(-G'( -t/s ))... derive: ds=-G dx, dl=(t - lam ds)/s
            dxc = solve(-G.T @ (t / s)); dsc = -G @ dxc; dlc = (t - lam * dsc) / s
            dx2, ds2, dl2 = dx + dxc, ds + dsc, dl + dlc
            am2 = E.alpha_max(s, ds2, lam, dl2)
            ncorr += 1
            if am2 >= am + gam * (at - am) or am2 >= 0.999:
                dx, ds, dl, am = dx2, ds2, dl2, am2
            else:
                break
        tau = max(0.99, 1.0 - ratio)
        a = min(1.0, tau * am)
        x, s, lam = x + a * dx, s + a * ds, lam + a * dl
    return x, it, lam, 1, ncorr

def run(o, label, cold=False, cc=0.4):
    its = np.zeros((E.T, E.N)); cor = np.zeros((E.T, E.N)); err = 0; bad = 0
    wsx = [None] * E.N
    for rec in E.data:
        i, k = rec["i"], rec["k"]
        G, h = rec["G"][:-1], rec["h"][:-1]
        x, it, lam, st, nc = ipm_g(rec["H"], rec["q"], G, h, ws=None if cold else wsx[i], o=o)
        its[k, i] = it; cor[k, i] = nc; bad += st
        err = max(err, np.abs(x - rec["x"]).max() / (1 + np.abs(rec["x"]).max()))
        wsx[i] = (x, lam) if (lam is not None and st == 0) else None
    a = its[5:]; c = a + cc * cor[5:]
    print("%-46s its mean %.2f p99 %d max %d permax %.1f | cost mean %.2f p99 %.1f max %.1f permax %.1f | err %.1e bad %d" % (label, a.mean(), np.percentile(a, 99), a.max(), a.max(axis=1).mean(), c.mean(), np.percentile(c, 99), c.max(), c.max(axis=1).mean(), err, bad), flush=True)

if __name__ == "__main__":
    run(dict(K=0), "K=0 (current)")
    run(dict(K=1), "K=1 delta .3")
    run(dict(K=2), "K=2 delta .3")
    run(dict(K=2, delta=0.5), "K=2 delta .5")
    run(dict(K=3, delta=0.3), "K=3 delta .3")
    run(dict(K=2, mutfloor=1e-3), "K=2 mutfloor 1e-3")
    run(dict(K=0, cold="viol"), "K=0 cold viol")
    run(dict(K=2, cold="viol"), "K=2 cold viol")
