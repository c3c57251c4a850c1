"""Exact dense convex QP solver used by the oracle (test infrastructure, see oracle/__init__.py).

Solves   min 1/2 z'Hz + q'z   s.t.  A z <= b,  lb <= z <= ub
which is the problem the reference hands to JuMP in ``init_optimization!``
(reference src/controller/linmpc.jl:323-339: ``@variable Zmin<=Zvar<=Zmax``,
``@constraint A[i_b,:]*Zvar .<= b[i_b]``, ``@objective Min obj_quadprog(Zvar,H,q)``,
``obj_quadprog`` = src/general.jl:107).

The reference's default backend is the third-party OSQP ADMM solver (eps 1e-3); exact
backends (DAQP, Ipopt) are asserted equal to 1e-10 / 1e-3 in the reference's tests
(test/5_test_extensions.jl:33,41).  The oracle therefore returns the EXACT optimum:
a Mehrotra interior-point solve followed by an active-set polish that solves the KKT
system of the identified active set; the returned ``kkt`` residual is the certificate.
The product's CUDA solver uses a different formulation (input-level coordinates, merged
rows, its own IPM) so agreement between the two is a genuine cross-check.
"""
from __future__ import annotations

import numpy as np

OPTIMAL, ITERATION_LIMIT, INFEASIBLE = 0, 1, 2


def _stack_constraints(n, A, b, lb, ub):
    rows, rhs = [], []
    if A is not None and len(b):
        rows.append(np.asarray(A, float).reshape(-1, n))
        rhs.append(np.asarray(b, float))
    if lb is not None:
        idx = np.flatnonzero(np.isfinite(lb))
        if idx.size:
            G = np.zeros((idx.size, n))
            G[np.arange(idx.size), idx] = -1.0
            rows.append(G)
            rhs.append(-np.asarray(lb, float)[idx])
    if ub is not None:
        idx = np.flatnonzero(np.isfinite(ub))
        if idx.size:
            G = np.zeros((idx.size, n))
            G[np.arange(idx.size), idx] = 1.0
            rows.append(G)
            rhs.append(np.asarray(ub, float)[idx])
    if not rows:
        return np.zeros((0, n)), np.zeros(0)
    return np.vstack(rows), np.concatenate(rhs)


def _sym_solve(M, r):
    try:
        L = np.linalg.cholesky(M)
        return np.linalg.solve(L.T, np.linalg.solve(L, r))
    except np.linalg.LinAlgError:
        return np.linalg.lstsq(M, r, rcond=None)[0]


def ipm(H, q, G, h, max_iter=200, tol=1e-11):
    """Mehrotra predictor-corrector on  min 1/2z'Hz+q'z  s.t. Gz+s=h, s>=0."""
    n, m = H.shape[0], G.shape[0]
    if m == 0:
        return _sym_solve(H, -q), np.zeros(0), np.zeros(0), 0, True
    z = _sym_solve(H + 1e-9 * np.eye(n) * max(1.0, np.abs(np.diag(H)).max()), -q)
    s = h - G @ z
    scale = max(1.0, np.abs(h).max(), np.abs(s).max())
    s = np.maximum(s, 1e-2 * scale)
    lam = np.ones(m) * max(1.0, np.abs(q).max()) / scale * 1e-2 + 1e-8
    ok = False
    it = 0
    for it in range(1, max_iter + 1):
        rd = H @ z + q + G.T @ lam
        rp = G @ z + s - h
        mu = s @ lam / m
        if (np.abs(rd).max() <= tol * (1 + np.abs(q).max())
                and np.abs(rp).max() <= tol * (1 + np.abs(h).max())
                and mu <= tol * (1 + abs(0.5 * z @ H @ z + q @ z))):
            ok = True
            break
        d = lam / s
        Phi = H + G.T @ (d[:, None] * G)
        Phi += 1e-14 * np.eye(n) * max(1.0, np.abs(np.diag(Phi)).max())
        # affine (predictor)
        rhs = -rd - G.T @ (d * rp - lam)
        dz = _sym_solve(Phi, rhs)
        ds = -rp - G @ dz
        dl = -lam - d * ds
        a = _steplen(s, ds, lam, dl)
        mu_a = (s + a * ds) @ (lam + a * dl) / m
        sig = (mu_a / mu) ** 3 if mu > 0 else 0.0
        # corrector
        rc = s * lam + ds * dl - sig * mu
        rhs = -rd - G.T @ ((lam * rp - rc) / s)
        dz = _sym_solve(Phi, rhs)
        ds = -rp - G @ dz
        dl = -(rc + lam * ds) / s
        a = min(1.0, 0.995 * _steplen(s, ds, lam, dl))
        z = z + a * dz
        s = s + a * ds
        lam = lam + a * dl
        if not np.all(np.isfinite(z)):
            break
    return z, s, lam, it, ok


def _steplen(s, ds, lam, dl):
    a = 1.0
    neg = ds < 0
    if neg.any():
        a = min(a, (-s[neg] / ds[neg]).min())
    neg = dl < 0
    if neg.any():
        a = min(a, (-lam[neg] / dl[neg]).min())
    return a


def kkt_residual(H, q, G, h, z, lam):
    """max(stationarity, primal infeasibility, dual infeasibility, complementarity), scaled."""
    if G.shape[0] == 0:
        return float(np.abs(H @ z + q).max() / (1 + np.abs(q).max()))
    rd = np.abs(H @ z + q + G.T @ lam).max() / (1 + np.abs(q).max())
    slack = h - G @ z
    rp = max(0.0, (-slack).max()) / (1 + np.abs(h).max())
    rl = max(0.0, (-lam).max()) / (1 + np.abs(lam).max())
    rc = np.abs(lam * slack).max() / (1 + np.abs(lam).max() * (1 + np.abs(h).max()))
    return float(max(rd, rp, rl, rc))


def _polish(H, q, G, h, z, lam, max_iter=100):
    """Active-set polish: solve the equality-constrained KKT system of the guessed active set,
    add the most violated / drop the most negative-multiplier row until KKT holds."""
    n, m = H.shape[0], G.shape[0]
    slack = h - G @ z
    hs = 1 + np.abs(h)
    active = (slack < 1e-6 * hs) & (lam > 1e-9 * (1 + np.abs(lam).max()))
    best = (kkt_residual(H, q, G, h, z, lam), z, lam)
    for _ in range(max_iter):
        idx = np.flatnonzero(active)
        k = idx.size
        if k:
            Ga = G[idx]
            KKT = np.block([[H, Ga.T], [Ga, np.zeros((k, k))]])
            rhs = np.concatenate([-q, h[idx]])
            sol = np.linalg.lstsq(KKT, rhs, rcond=None)[0]
            zc, la = sol[:n], sol[n:]
        else:
            zc, la = _sym_solve(H, -q), np.zeros(0)
        lc = np.zeros(m)
        lc[idx] = la
        res = kkt_residual(H, q, G, h, zc, lc)
        if res < best[0]:
            best = (res, zc, lc)
        viol = (G @ zc - h) / hs
        viol[idx] = 0.0
        worst = int(np.argmax(viol)) if m else -1
        changed = False
        if m and viol[worst] > 1e-11:
            active[worst] = True
            changed = True
        elif k and la.min() < -1e-11 * (1 + np.abs(la).max()):
            active[idx[int(np.argmin(la))]] = False
            changed = True
        if not changed:
            break
    return best


def solve_qp(H, q, A=None, b=None, lb=None, ub=None):
    """Return dict(z, lam, status, kkt, iters, J0) with the exact optimum of the dense QP.

    status: OPTIMAL (0) | ITERATION_LIMIT (1) | INFEASIBLE (2), mirroring the reference's
    status policy in ``optim_objective!`` (src/controller/execute.jl:482-503,
    src/general.jl:45-61): error statuses make the controller reuse the shifted previous
    solution; non-optimal-but-not-error statuses keep the solver's iterate.
    """
    H = np.asarray(H, float)
    H = 0.5 * (H + H.T)
    q = np.asarray(q, float)
    n = q.size
    G, h = _stack_constraints(n, A, b, lb, ub)
    z, s, lam, iters, ok = ipm(H, q, G, h)
    if not np.all(np.isfinite(z)):
        return dict(z=np.full(n, np.nan), lam=None, status=INFEASIBLE, kkt=np.inf, iters=iters)
    if G.shape[0]:
        res, z, lam = _polish(H, q, G, h, z, lam)
    else:
        res = kkt_residual(H, q, G, h, z, lam)
    if res <= 1e-8:
        status = OPTIMAL
    else:
        # primal infeasibility of the best point decides between the two non-optimal statuses
        viol = (G @ z - h).max() / (1 + np.abs(h).max()) if G.shape[0] else 0.0
        status = INFEASIBLE if viol > 1e-6 else ITERATION_LIMIT
    return dict(z=z, lam=lam, status=status, kkt=res, iters=iters,
                J0=float(0.5 * z @ H @ z + q @ z))
