import sys, pickle, time
import numpy as np
data = pickle.load(open("/tmp/mpcdata/qps_256_60.pkl", "rb"))
N = 256; T = 60

def alpha_max(s, ds, lam, dl):
    a = 1.0
    neg = ds < 0
    if neg.any(): a = min(a, (-s[neg] / ds[neg]).min())
    neg = dl < 0
    if neg.any(): a = min(a, (-lam[neg] / dl[neg]).min())
    return a

def gchol(A):
    try:
        return np.linalg.cholesky(A)
    except np.linalg.LinAlgError:
        pass
    A = A.copy(); n = A.shape[0]; L = np.zeros_like(A)
    for k in range(n):
        dk = A[k, k]
        if not dk > 1e-280: dk = 1e200
        L[k, k] = np.sqrt(dk)
        L[k+1:, k] = A[k+1:, k] / L[k, k]
        A[k+1:, k+1:] -= np.outer(L[k+1:, k], L[k+1:, k])
    return L

def ipm(H, q, G, h, ws=None, opt=None, tol=1e-11, max_iter=50):
    o = dict(tau=0.99, adapt_tau=False, s_floor=1e-2, lmin=1e-4, center=None, sigpow=3, tolmu_fac=1e-3, gondzio=0)
    if opt: o.update(opt)
    n, m = q.size, h.size
    nz = n - 1
    x = np.zeros(n)
    L = np.linalg.cholesky(H[:nz, :nz]); x[:nz] = -np.linalg.solve(L.T, np.linalg.solve(L, q[:nz]))
    slack0 = h - G @ x
    hscale = 1.0 + np.abs(h).max()
    if slack0.min() >= -1e-12 * hscale: return x, 0, None, 0
    qs = 1.0 + np.abs(q).max()
    mu0 = max(1e-2 * qs * hscale / m, 1e-8)
    s = np.maximum(slack0, 1e-2 * hscale); lam = mu0 / s
    if ws is not None:
        xw, lw = ws
        x = xw.copy()
        sl = h - G @ x
        if o["center"] is None:
            s = np.maximum(sl, o["s_floor"] * hscale)
            lam = np.maximum(lw, o["lmin"] * qs / hscale)
        else:
            mut = o["center"] * qs * hscale / m
            lamf = np.maximum(lw, 1e-300)
            s = np.maximum(sl, np.minimum(mut / lamf, o["s_floor"] * hscale))
            lam = np.maximum(lw, mut / s)
    best = 1e300
    tol_mu = o["tolmu_fac"] * tol
    for it in range(max_iter + 1):
        Hxq = H @ x + q; Gl = G.T @ lam
        rd = Hxq + Gl; rp = G @ x + s - h; mu = s @ lam / m
        qd = qs + max(np.abs(Hxq).max(), np.abs(Gl).max())
        merit = max(np.abs(rd).max() / (tol * qd), np.abs(rp).max() / (tol * hscale), mu * m / (tol_mu * qs * hscale))
        if merit <= 1.0 or (best <= 1e3 and merit >= best): return x, it, lam, 0
        best = min(best, merit)
        if it == max_iter: return x, it, lam, 1
        d = lam / s
        Phi = H + G.T @ (d[:, None] * G)
        Lc = gchol(Phi)
        solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        dx = solve(-rd - G.T @ (d * rp - lam)); ds = -rp - G @ dx; dl = -lam - d * ds
        a = alpha_max(s, ds, lam, dl)
        mu_a = (s + a * ds) @ (lam + a * dl) / m
        sig = (mu_a / mu) ** o["sigpow"]
        rc = s * lam + ds * dl - sig * mu
        dx = solve(-rd - G.T @ ((lam * rp - rc) / s)); ds = -rp - G @ dx; dl = -(rc + lam * ds) / s
        am = alpha_max(s, ds, lam, dl)
        tau = o["tau"]
        if o["adapt_tau"]:
            tau = max(o["tau"], 1.0 - o["adapt_tau"] * mu_a / mu)  # close to 1 when the affine step nearly kills the gap
        a = min(1.0, tau * am)
        x, s, lam = x + a * dx, s + a * ds, lam + a * dl
    return x, it, lam, 1

def run(opt, label, nmax=N):
    its = np.zeros((T, nmax), int); err = 0.0; bad = 0
    wsx = [None] * nmax
    t0 = time.time()
    for rec in data:
        i, k = rec["i"], rec["k"]
        if i >= nmax: continue
        x, it, lam, st = ipm(rec["H"], rec["q"], rec["G"], rec["h"], ws=wsx[i], opt=opt)
        its[k, i] = it; bad += st
        e = np.abs(x - rec["x"]).max() / (1 + np.abs(rec["x"]).max()); err = max(err, e)
        wsx[i] = (x, lam) if (lam is not None and st == 0) else None
    a = its[5:]
    print("%-40s mean %.2f p90 %d p99 %d max %d  permax-mean %.1f  err %.1e bad %d  (%.0fs)" % (label, a.mean(), np.percentile(a, 90), np.percentile(a, 99), a.max(), a.max(axis=1).mean(), err, bad, time.time() - t0), flush=True)
    return its

if __name__ == "__main__":
    run(None, "baseline")
    run(dict(adapt_tau=1.0), "adapt tau 1-mu_a/mu")
    run(dict(tau=0.995), "tau .995")
    run(dict(tau=0.999), "tau .999")
    run(dict(center=1e-2), "center 1e-2")
    run(dict(center=1e-4), "center 1e-4")
    run(dict(center=1e-6), "center 1e-6")
