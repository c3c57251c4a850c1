import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import numpy as np
import ipm_exp as E
rec = E.data[100]
H, q = rec["H"], rec["q"]; G, h = rec["G"][:-1].copy(), rec["h"][:-1].copy()
n = q.size
for gap in (0.5, 0.01):
    g1 = np.zeros(n); g1[0] = 1; g2 = np.zeros(n); g2[0] = -1
    G2 = np.vstack([G, g1, g2]); h2 = np.concatenate([h, [-1.0, 1.0 - gap]])   # x0 <= -1 and x0 >= 1-gap... infeasible: -x0 <= 1-gap-> x0 >= gap-1 ; with x0<=-1 infeasible by `gap`
    m = h2.size; nz = n - 1
    x = np.zeros(n); L = np.linalg.cholesky(H[:nz, :nz]); x[:nz] = -np.linalg.solve(L.T, np.linalg.solve(L, q[:nz]))
    hscale = 1 + np.abs(h2).max(); qs = 1 + np.abs(q).max(); mu0 = max(1e-2 * qs * hscale / m, 1e-8)
    s = np.maximum(h2 - G2 @ x, 1e-2 * hscale); lam = mu0 / s
    print("gap", gap)
    for it in range(40):
        rd = H @ x + q + G2.T @ lam; rp = G2 @ x + s - h2; mu = s @ lam / m
        d = lam / s; Lc = E.gchol(H + G2.T @ (d[:, None] * G2)); solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        dx = solve(-rd - G2.T @ (d * rp - lam)); ds = -rp - G2 @ dx; dl = -lam - d * ds
        aa = E.alpha_max(s, ds, lam, dl); mua = (s + aa * ds) @ (lam + aa * dl) / m; sig = (mua / mu) ** 3
        rc = s * lam + ds * dl - sig * mu
        dx = solve(-rd - G2.T @ ((lam * rp - rc) / s)); ds = -rp - G2 @ dx; dl = -(rc + lam * ds) / s
        am = E.alpha_max(s, ds, lam, dl); tau = min(max(0.99, 1 - mua / mu), 1 - 1e-6); a = min(1, tau * am)
        print("%2d rd %.1e rp %.2e mu %.1e a %.2e lam_max %.1e" % (it, np.abs(rd).max(), np.abs(rp).max(), mu, a, lam.max()))
        x, s, lam = x + a * dx, s + a * ds, lam + a * dl
