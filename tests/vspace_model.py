"""numpy model of the ALGORITHM the CUDA kernel runs (test helper, not product, not oracle).

The kernel does not solve the reference's QP in the reference's coordinates.  It works in
"input-level" coordinates  v_l = u0(k+j_l) - u0(k-1) = sum_{i<=l} DU_i  (one level per move
block and input), in which
  * every U bound is a bound on ONE variable (rows that repeat inside a move block merge),
  * every DU bound touches TWO variables (v_l - v_{l-1}),
  * only the Yhat / terminal-state rows stay dense (E_v = E*D, e_v = ex*D).
This file restates that transformation and the structured Mehrotra iteration so that the
formulation can be validated against oracle/ on the CPU, independently of CUDA indexing.
"""
import numpy as np


def dmat(nu, Hc):
    """z = D v  (DU_l = v_l - v_{l-1})."""
    nz = nu * Hc
    D = np.eye(nz)
    for i in range(nu, nz):
        D[i, i - nu] = -1.0
    return D


def compile_rows(mpc):
    """Build the merged row table from an oracle LinMPC: returns dict with
    sparse rows (i1, i2, sigma, c, hbase, shift_ch) and dense rows (base t, sigma, c, kind)."""
    c = mpc.con
    nu, Hp, Hc, nb = mpc.model.nu, mpc.Hp, mpc.Hc, mpc.nb
    block_of_t = np.repeat(np.arange(Hc), nb)
    hard = lambda cc: (mpc.neps == 0) or cc == 0.0
    sparse = {}

    def add(i1, i2, sigma, cc, hval, ch):
        key = (i1, i2, sigma, cc, ch)
        sparse[key] = min(sparse.get(key, np.inf), hval)
    for t in range(Hp):
        for ch in range(nu):
            k = t * nu + ch
            var = block_of_t[t] * nu + ch
            if np.isfinite(c.U0min[k]):
                add(var, -1, -1, c.C_umin[k] if mpc.neps else 0.0, -c.U0min[k], ch)
            if np.isfinite(c.U0max[k]):
                add(var, -1, +1, c.C_umax[k] if mpc.neps else 0.0, c.U0max[k], ch)
    for k in range(nu * Hc):
        i2 = k - nu if k >= nu else -1
        if np.isfinite(c.DUmin[k]):
            add(k, i2, -1, c.C_dumin[k] if mpc.neps else 0.0, -c.DUmin[k], -1)
        if np.isfinite(c.DUmax[k]):
            add(k, i2, +1, c.C_dumax[k] if mpc.neps else 0.0, c.DUmax[k], -1)
    srows = [dict(i1=k[0], i2=k[1], sigma=k[2], c=k[3], ch=k[4], hbase=v) for k, v in sparse.items()]
    drows = []
    for t in range(mpc.model.ny * Hp):
        if np.isfinite(c.Y0min[t]):
            drows.append(dict(kind="y", t=t, sigma=-1, c=c.C_ymin[t] if mpc.neps else 0.0))
        if np.isfinite(c.Y0max[t]):
            drows.append(dict(kind="y", t=t, sigma=+1, c=c.C_ymax[t] if mpc.neps else 0.0))
    for i in range(mpc.estim.nxhat):
        if np.isfinite(c.xhat0min[i]):
            drows.append(dict(kind="x", t=i, sigma=-1, c=c.c_xmin[i] if mpc.neps else 0.0))
        if np.isfinite(c.xhat0max[i]):
            drows.append(dict(kind="x", t=i, sigma=+1, c=c.c_xmax[i] if mpc.neps else 0.0))
    return srows, drows


def build_qp_v(mpc):
    """Dense (G, h, H, q) of the v-space problem for the CURRENT step of an oracle LinMPC on
    which initpred/linconstraint were already run.  x = [v; eps]."""
    nu, Hc, neps = mpc.model.nu, mpc.Hc, mpc.neps
    nz = nu * Hc
    n = nz + neps
    D = dmat(nu, Hc)
    Dt = np.eye(n)
    Dt[:nz, :nz] = D
    H = Dt.T @ mpc.Htilde @ Dt
    q = Dt.T @ mpc.qtilde
    Ev, exv = mpc.E @ D, mpc.ex @ D
    srows, drows = compile_rows(mpc)
    G, h = [], []
    for r in srows:
        g = np.zeros(n)
        g[r["i1"]] = r["sigma"]
        if r["i2"] >= 0:
            g[r["i2"]] = -r["sigma"]
        if neps:
            g[-1] = -r["c"]
        G.append(g)
        h.append(r["hbase"] - (r["sigma"] * mpc.Tu_lastu0[r["ch"]] if r["ch"] >= 0 else 0.0))
    for r in drows:
        g = np.zeros(n)
        if r["kind"] == "y":
            g[:nz] = r["sigma"] * Ev[r["t"]]
            bound = mpc.con.Y0max[r["t"]] if r["sigma"] > 0 else mpc.con.Y0min[r["t"]]
            h.append(r["sigma"] * (bound - mpc.F[r["t"]]))
        else:
            g[:nz] = r["sigma"] * exv[r["t"]]
            bound = mpc.con.xhat0max[r["t"]] if r["sigma"] > 0 else mpc.con.xhat0min[r["t"]]
            h.append(r["sigma"] * (bound - mpc.con.fx[r["t"]]))
        if neps:
            g[-1] = -r["c"]
        G.append(g)
    # the reference's eps >= 0 row is not compiled: all softness weights are >= 0 (construct.jl:456-506), so
    # eps < 0 is never optimal and the row is redundant (its vanishing multiplier only slows the IPM down)
    G = np.array(G).reshape(-1, n)
    return H, q, G, np.array(h), Dt


def ipm_device_model(H, q, G, h, max_iter=60, tol=1e-9, neps=1, verbose=False, tol_mu=1e-12):
    """The kernel's solver: unconstrained shortcut, then Mehrotra predictor-corrector with a
    single step length, cold start s = max(h - Gx0, smin), lambda = mu0 / s."""
    n, m = q.size, h.size
    nz = n - neps
    x = np.zeros(n)
    try:
        L = np.linalg.cholesky(H[:nz, :nz])
        x[:nz] = -np.linalg.solve(L.T, np.linalg.solve(L, q[:nz]))
    except np.linalg.LinAlgError:
        pass
    slack0 = h - G @ x
    hscale = 1.0 + np.abs(h).max() if m else 1.0
    if m == 0 or slack0.min() >= -1e-12 * hscale:
        return x, 0, 0
    qs = 1.0 + np.abs(q).max()
    s = np.maximum(slack0, 1e-2 * hscale)
    viol = max(0.0, -(slack0).min())
    if neps:  # start the slack variable where the soft rows are (nearly) satisfied
        pass
    lam = np.full(m, 1.0)
    mu0 = max(1e-2 * qs * hscale / m, 1e-8)
    lam = mu0 / s
    status = 1
    it = 0
    for it in range(1, max_iter + 1):
        rd = H @ x + q + G.T @ lam
        rp = G @ x + s - h
        mu = s @ lam / m
        if verbose:
            print(it, np.abs(rd).max(), np.abs(rp).max(), mu)
        qd = qs + max(np.abs(H @ x + q).max(), np.abs(G.T @ lam).max())
        if np.abs(rd).max() <= tol * qd and np.abs(rp).max() <= tol * hscale and mu * m <= tol_mu * qs * hscale:
            status = 0
            it -= 1
            break
        d = lam / s
        Phi = H + G.T @ (d[:, None] * G)
        Lc = np.linalg.cholesky(Phi + 1e-300 * np.eye(n))
        solve = lambda r: np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
        dx = solve(-rd - G.T @ (d * rp - lam))
        ds = -rp - G @ dx
        dl = -lam - d * ds
        a = _alpha(s, ds, lam, dl)
        mu_a = (s + a * ds) @ (lam + a * dl) / m
        sig = (mu_a / mu) ** 3
        rc = s * lam + ds * dl - sig * mu
        dx = solve(-rd - G.T @ ((lam * rp - rc) / s))
        ds = -rp - G @ dx
        dl = -(rc + lam * ds) / s
        tau = min(max(0.99, 1.0 - mu_a / mu), 1.0 - 1e-6)  # fraction to the boundary -> 1 as the affine step closes the gap
        a = min(1.0, tau * _alpha(s, ds, lam, dl))
        x, s, lam = x + a * dx, s + a * ds, lam + a * dl
        if not np.isfinite(x).all():
            status = 2
            break
    if status == 1:
        rp = G @ x + s - h
        if np.abs(rp).max() > 1e-6 * hscale:
            status = 2
    return x, it, status


def _alpha(s, ds, lam, dl):
    a = 1.0
    neg = ds < 0
    if neg.any():
        a = min(a, (-s[neg] / ds[neg]).min())
    neg = dl < 0
    if neg.any():
        a = min(a, (-lam[neg] / dl[neg]).min())
    return a
