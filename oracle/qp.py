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
    best = (np.inf, z, s, lam)
    for it in range(1, max_iter + 1):
        rd = H @ z + q + G.T @ lam
        rp = G @ z + s - h
        mu = s @ lam / m
        e_d = np.abs(rd).max() / (1 + np.abs(q).max())
        e_p = np.abs(rp).max() / (1 + np.abs(h).max())
        e_c = mu / (1 + abs(0.5 * z @ H @ z + q @ z))
        merit = max(e_d, e_p, e_c)
        if merit < best[0]:
            best = (merit, z, s, lam)
        if merit <= tol:
            ok = True
            break
        if mu < 1e-30 or s.min() < 1e-150:  # converged as far as fp64 allows
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
        if not (np.all(np.isfinite(z)) and np.all(np.isfinite(lam)) and np.all(np.isfinite(s))):
            break
    _, z, s, lam = best
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


def _kkt_solve(H, N, q, hA):
    """Solve [H N; N' 0][x; u] = [-q; hA] (LU + two refinement steps; least squares if singular)."""
    n, k = H.shape[0], N.shape[1]
    KKT = np.block([[H, N], [N.T, np.zeros((k, k))]])
    rhs = np.concatenate([-q, hA])
    try:
        import scipy.linalg as sla
        lu = sla.lu_factor(KKT)
        sol = sla.lu_solve(lu, rhs)
        for _ in range(2):
            sol = sol + sla.lu_solve(lu, rhs - KKT @ sol)
        if not np.all(np.isfinite(sol)):
            raise np.linalg.LinAlgError
    except Exception:
        sol = np.linalg.lstsq(KKT, rhs, rcond=None)[0]
    return sol[:n], sol[n:]


def goldfarb_idnani(H, q, G, h, max_iter=2000, ptol=1e-12):
    """Dual active-set method of Goldfarb & Idnani (1983) for strictly convex QPs,
    min 1/2 z'Hz + q'z  s.t. Gz <= h: exact (finite termination), the algorithm family DAQP --
    the solver the reference's 1e-10 cross-checks use (test/5_test_extensions.jl:33) -- belongs to.
    Dense re-solves instead of factor updates: the problems are small.  Returns (z, lam, ok)."""
    n, m = H.shape[0], G.shape[0]
    L = np.linalg.cholesky(H)
    Hinv = lambda r: np.linalg.solve(L.T, np.linalg.solve(L, r))
    x = Hinv(-q)
    act, u = [], []
    refits = 0
    hs = 1.0 + np.abs(h)
    for _ in range(max_iter):
        viol = (G @ x - h) / hs
        if act:
            viol[act] = -np.inf
        p = int(np.argmax(viol)) if m else -1
        if m == 0 or viol[p] <= ptol:
            if act and refits < 8:
                # remove the drift of the incremental updates: re-solve the KKT system of the active set
                refits += 1
                x, unew = _kkt_solve(H, G[act].T, q, h[act])
                keep = unew > 0
                act = [a for a, kp in zip(act, keep) if kp]
                u = [float(v) for v, kp in zip(unew, keep) if kp]
                viol = (G @ x - h) / hs
                if act:
                    viol[act] = -np.inf
                if keep.all() and viol.max() <= ptol:
                    lam = np.zeros(m)
                    lam[act] = u
                    return x, lam, True
                if not keep.all():
                    # dropped rows: restore stationarity for the reduced set before continuing
                    if act:
                        x, unew = _kkt_solve(H, G[act].T, q, h[act])
                        u = [float(v) for v in unew]
                    else:
                        x = Hinv(-q)
                continue
            lam = np.zeros(m)
            lam[act] = u
            return x, lam, True
        nrm = G[p]
        up = 0.0
        while True:
            if act:
                N = G[act].T
                HiN = Hinv(N)
                S = N.T @ HiN
                Hin = Hinv(nrm)
                r = np.linalg.lstsq(S, N.T @ Hin, rcond=None)[0]
                z = Hin - HiN @ r
            else:
                r = np.zeros(0)
                z = Hinv(nrm)
            nz = float(nrm @ z)
            # n is (numerically) in the span of the active normals: relative to its H^-1 norm
            znull = nz <= 1e-10 * float(nrm @ Hinv(nrm))
            t1, k = np.inf, -1
            for j in range(len(act)):
                if r[j] > 1e-14 * (1 + np.abs(r).max()):
                    tj = u[j] / r[j]
                    if tj < t1:
                        t1, k = tj, j
            t2 = np.inf if znull else float(G[p] @ x - h[p]) / nz
            t = min(t1, t2)
            if not np.isfinite(t):
                return x, None, False  # infeasible
            if act:
                u = list(np.asarray(u) - t * r)
            up += t
            if not znull:
                x = x - t * z
            if t == t2 and not znull:
                act.append(p)
                u.append(up)
                break
            del act[k]
            del u[k]
    return x, None, False


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
            zc, la = _kkt_solve(H, G[idx].T, q, h[idx])
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
    # 1) exact dual active-set solve when H is safely positive definite
    try:
        ev = np.linalg.eigvalsh(H)
        if ev[0] > 1e-10 * max(1.0, ev[-1]):
            zg, lg, okg = goldfarb_idnani(H, q, G, h)
            if okg:
                res = kkt_residual(H, q, G, h, zg, lg)
                if res <= 1e-9:
                    return dict(z=zg, lam=lg, status=OPTIMAL, kkt=res, iters=0,
                                J0=float(0.5 * zg @ H @ zg + q @ zg))
    except np.linalg.LinAlgError:
        pass
    # 2) interior point + active-set polish (semidefinite H, or infeasible problems)
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


def solve_qp_eq(H, q, A, b, Aeq, beq, lb=None, ub=None):
    """Exact optimum of  min 1/2 z'Hz + q'z  s.t.  A z <= b,  Aeq z = beq,  lb <= z <= ub  (H only positive
    semidefinite on the whole space, as the MultipleShooting Hessian is).  The equalities are eliminated with a GENERIC
    orthonormal null-space basis of Aeq (SVD; nothing of the problem's structure is used): z = z_p + N w, then the
    reduced inequality QP goes through ``solve_qp``.  Returns the same dict, with z in the original space."""
    H = np.asarray(H, float)
    H = 0.5 * (H + H.T)
    q = np.asarray(q, float)
    n = q.size
    Aeq = np.asarray(Aeq, float).reshape(-1, n)
    beq = np.asarray(beq, float).reshape(-1)
    G, h = _stack_constraints(n, A, b, lb, ub)
    if Aeq.shape[0] == 0:
        return solve_qp(H, q, G, h)
    U, s, Vt = np.linalg.svd(Aeq, full_matrices=True)
    rank = int((s > 1e-12 * max(Aeq.shape) * (s[0] if s.size else 1.0)).sum())
    zp = Vt[:rank].T @ ((U[:, :rank].T @ beq) / s[:rank])
    if np.abs(Aeq @ zp - beq).max() > 1e-8 * (1 + np.abs(beq).max()):
        return dict(z=np.full(n, np.nan), lam=None, status=INFEASIBLE, kkt=np.inf, iters=0)
    N = Vt[rank:].T
    Hr = N.T @ H @ N
    qr = N.T @ (H @ zp + q)
    sol = solve_qp(Hr, qr, G @ N, h - G @ zp)
    z = zp + N @ sol["z"] if np.all(np.isfinite(sol["z"])) else np.full(n, np.nan)
    out = dict(sol)
    out["z"] = z
    out["J0"] = float(0.5 * z @ H @ z + q @ z) if np.all(np.isfinite(z)) else np.nan
    return out
