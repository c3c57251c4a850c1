import sys; _H = __import__("os").path.dirname(__import__("os").path.abspath(__file__)); sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(_H))); sys.path.insert(0, _H)
import numpy as np
from oracle.linmpc import LinModel as OLinModel
from oracle.mhe import MovingHorizonEstimator as OMHE
from oracle import qp as _qp
import ipm_exp as E
rng = np.random.default_rng(3)
nx, nu, ny, He = 4, 2, 2, 8
found = []
for trial in range(12):
    A = rng.standard_normal((nx, nx)); A *= rng.uniform(0.5, 0.9) / np.abs(np.linalg.eigvals(A)).max()
    m = OLinModel(A, rng.standard_normal((nx, nu)), rng.standard_normal((ny, nx)))
    o = OMHE(m, He=He, nint_ym=0, Cwt=np.inf).setconstraint(xhatmin=[-10] * nx, xhatmax=[10] * nx, whatmin=[-0.5] * nx, whatmax=[0.5] * nx, vhatmin=[-2.0] * ny, vhatmax=[2.0] * ny)
    x = np.zeros(nx)
    for k in range(30):
        u = rng.choice([-1.0, 1.0], nu)
        x = A @ x + m.Bu @ u + rng.standard_normal(nx) / nx
        y = m.C @ x + rng.standard_normal(ny)
        o.preparestate(y)
        if o.last_qp["status"] == _qp.INFEASIBLE:
            P = o.build_qp(); found.append((P["H"], P["q"], P["A"], P["b"]))
        o.updatestate(u, y)
print("infeasible QPs found", len(found))
def trace(H, q, G, h, maxit=50):
    n, m = q.size, h.size
    x = -np.linalg.solve(H, q)
    hscale = 1 + np.abs(h).max(); qs = 1 + np.abs(q).max(); mu0 = max(1e-2 * qs * hscale / m, 1e-8)
    s = np.maximum(h - G @ x, 1e-2 * hscale); lam = mu0 / s
    out = []
    stall = 0
    for it in range(maxit):
        Gl = G.T @ lam; rd = H @ x + q + Gl; rp = G @ x + s - h; mu = s @ lam / m
        glmax = np.abs(Gl).max(); lmax = lam.max(); hl = h @ lam
        cert = glmax <= 1e-7 * lmax * (1 + 0) and hl <= -1e-7 * lmax * hscale
        d = lam / s; Lc = E.gchol(H + G.T @ (d[:, None] * G)); solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        dx = solve(-rd - G.T @ (d * rp - lam)); ds = -rp - G @ dx; dl = -lam - d * ds
        aa = E.alpha_max(s, ds, lam, dl); mua = (s + aa * ds) @ (lam + aa * dl) / m; sig = (mua / mu) ** 3
        rc = s * lam + ds * dl - sig * mu
        dx = solve(-rd - G.T @ ((lam * rp - rc) / s)); ds = -rp - G @ dx; dl = -(rc + lam * ds) / s
        tau = min(max(0.99, 1 - mua / mu), 1 - 1e-6); a = min(1, tau * E.alpha_max(s, ds, lam, dl))
        out.append((it, np.abs(rp).max(), mu, a, lmax, glmax / lmax, hl / (lmax * hscale), cert))
        stall = stall + 1 if (a < 1e-8 and np.abs(rp).max() > 1e-6 * hscale) else 0
        if stall >= 2: out.append(("collapse-exit", it)); break
        x, s, lam = x + a * dx, s + a * ds, lam + a * dl
        if not np.isfinite(lam).all(): out.append(("nonfinite", it)); break
    return out
for H, q, G, h in found[:4]:
    o = trace(H, q, G, h)
    first_cert = next((r[0] for r in o if len(r) == 8 and r[7]), None)
    print("n", q.size, "m", h.size, "first cert at", first_cert, "end:", o[-1] if len(o[-1]) == 2 else ("cap", o[-1][0]))
    for r in o[:40:3]:
        if len(r) == 8: print("   it %2d rp %.2e mu %.1e a %.1e lmax %.1e gl/l %.1e hl/(l h) %.1e cert %s" % r)
