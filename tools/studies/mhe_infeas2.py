"""Earlier Farkas-ray exit for slack-free problems: false positives / exit iteration of the checkpoint rule
(start, spacing) on MHE windows whose status is known from the exact oracle."""
import sys; _H = __import__("os").path.dirname(__import__("os").path.abspath(__file__)); sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(_H))); sys.path.insert(0, _H)
import numpy as np
from oracle.linmpc import LinModel as OLinModel
from oracle.mhe import MovingHorizonEstimator as OMHE
from oracle import qp as _qp
import ipm_exp as E
rng = np.random.default_rng(5)
nx, nu, ny, He = 4, 2, 2, 8
qps = []
for trial in range(40):
    A = rng.standard_normal((nx, nx)); A *= rng.uniform(0.5, 0.9) / np.abs(np.linalg.eigvals(A)).max()
    m = OLinModel(A, rng.standard_normal((nx, nu)), rng.standard_normal((ny, nx)))
    vb = rng.choice([2.0, 2.5, 3.0])
    o = OMHE(m, He=He, nint_ym=0, Cwt=np.inf).setconstraint(xhatmin=[-10] * nx, xhatmax=[10] * nx, whatmin=[-0.5] * nx, whatmax=[0.5] * nx, vhatmin=[-vb] * ny, vhatmax=[vb] * ny)
    x = np.zeros(nx)
    for k in range(40):
        u = rng.choice([-1.0, 1.0], nu)
        x = A @ x + m.Bu @ u + rng.standard_normal(nx) / nx
        y = m.C @ x + rng.standard_normal(ny)
        o.preparestate(y)
        P = o.build_qp()
        x0 = -np.linalg.solve(P["H"], P["q"])
        if (P["b"] - P["A"] @ x0).min() < -1e-12 * (1 + np.abs(P["b"]).max()):
            qps.append((P["H"], P["q"], P["A"], P["b"], o.last_qp["status"] == _qp.INFEASIBLE))
        o.updatestate(u, y)
print("active QPs", len(qps), "infeasible", sum(q[4] for q in qps), flush=True)

def run(H, q, G, h, start, step, need=2, mufac=100.0, maxit=50, tol=1e-11):
    n, m = q.size, h.size
    x = -np.linalg.solve(H, q)
    hscale = 1 + np.abs(h).max(); qs = 1 + np.abs(q).max(); mu0 = max(1e-2 * qs * hscale / m, 1e-8)
    s = np.maximum(h - G @ x, 1e-2 * hscale); lam = mu0 / s
    stall = 0; best = 1e300; mu_first = None; ep_chk = 1e300; nray = 0
    for it in range(maxit + 1):
        Hxq = H @ x + q; Gl = G.T @ lam; rd = Hxq + Gl; rp = G @ x + s - h; mu = s @ lam / m
        e_p = np.abs(rp).max(); qd = qs + max(np.abs(Hxq).max(), np.abs(Gl).max())
        merit = max(np.abs(rd).max() / (tol * qd), e_p / (tol * hscale), mu * m / (1e-3 * tol * qs * hscale))
        if merit <= 1 or (best <= 1e3 and merit >= best): return "opt", it
        best = min(best, merit)
        if it == maxit: return ("opt" if merit <= 1e3 else ("inf" if e_p > 1e-6 * hscale else "lim")), it
        d = lam / s; Lc = E.gchol(H + G.T @ (d[:, None] * G)); solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        dx = solve(-rd - G.T @ (d * rp - lam)); ds = -rp - G @ dx; dl = -lam - d * ds
        aa = E.alpha_max(s, ds, lam, dl); mua = (s + aa * ds) @ (lam + aa * dl) / m; sig = (mua / mu) ** 3
        rc = s * lam + ds * dl - sig * mu
        dx = solve(-rd - G.T @ ((lam * rp - rc) / s)); ds = -rp - G @ dx; dl = -(rc + lam * ds) / s
        tau = min(max(0.99, 1 - mua / mu), 1 - 1e-6); a = min(1, tau * E.alpha_max(s, ds, lam, dl))
        stall = stall + 1 if (a < 1e-8 and e_p > 1e-6 * hscale) else 0
        if stall >= 2: return "inf", it + 1
        if it == 0: mu_first = mu
        if it % step == 0:
            ray = it >= start and e_p > 0.5 * ep_chk and e_p > 1e-4 * hscale and mu > mufac * mu_first and h @ lam < 0
            nray = nray + 1 if ray else 0
            if nray >= need: return "inf", it + 1
            ep_chk = e_p
        x, s, lam = x + a * dx, s + a * ds, lam + a * dl
        if not np.isfinite(lam).all(): return "inf", it + 1
    return "lim", maxit

for start, step, need, mufac in [(16, 8, 2, 100.0), (8, 4, 2, 100.0), (12, 4, 2, 100.0), (8, 4, 3, 100.0), (8, 4, 2, 30.0), (6, 3, 3, 30.0), (8, 2, 4, 100.0)]:
    fp = fn = 0; it_inf = []; it_feas = []
    for H, q, G, h, inf in qps:
        st, it = run(H, q, G, h, start, step, need, mufac)
        if inf:
            it_inf.append(it); fn += st != "inf"
        else:
            it_feas.append(it); fp += st != "opt"
    print("start %2d step %d need %d mufac %5.0f: false positives %d / %d  missed %d / %d | mean exit it infeasible %.1f (max %d)  feasible mean %.1f max %d" % (
        start, step, need, mufac, fp, len(it_feas), fn, len(it_inf), np.mean(it_inf), max(it_inf), np.mean(it_feas), max(it_feas)), flush=True)
