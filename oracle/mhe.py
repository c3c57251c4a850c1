"""numpy fp64 restatement of the reference's linear ``MovingHorizonEstimator`` (TEST INFRASTRUCTURE ONLY).

LinModel + SingleShooting path of src/estimator/mhe/{construct,transcription,execute}.jl and the
time-varying ``KalmanFilter`` covariance recursion it uses for the arrival covariance
(src/estimator/kalman.jl:1235-1290).  ASCII names: x̂ -> xhat, Ŵ -> What, V̂ -> Vhat, P̄ -> Pbar,
Z̃ = [eps; xhat0_arr; What] (slack FIRST, construct.jl:1174-1178).
"""
from __future__ import annotations

import numpy as np

from . import qp as _qp
from .linmpc import StateEstimator


def init_predmat_mhe(He, A, Bu, Cm, Bd, Ddm, f_minus_x, direct):
    """src/estimator/mhe/transcription.jl:151-260 (E, G, J, B, exbar, EX, GX, JX, BX)."""
    nx, nu, nd, nym = A.shape[0], Bu.shape[1], Bd.shape[1], Cm.shape[0]
    nw = nx
    p = 0 if direct else 1
    Apow = [np.eye(nx)]
    for _ in range(He):
        Apow.append(Apow[-1] @ A)
    nCA = [-Cm @ Ap for Ap in Apow]                       # -Cm A^j, j = 0..He
    nCA_v = np.vstack(nCA)
    E = np.zeros((nym * He, nx + nw * He))
    col_begin, col_end = (1, He) if p == 0 else (0, He - 1)
    i = 0
    for j in range(col_begin, col_end + 1):
        rows = slice(i * nym, nym * He)
        E[rows, j * nw:(j + 1) * nw] = nCA_v[:nym * He - i * nym]
        i += 1
    if p == 0:
        E[:, :nx] = nCA_v[nym:]
    exbar = np.hstack([-np.eye(nx), np.zeros((nx, nw * He))])
    Apow_v = np.vstack(Apow)
    EX = np.zeros((nx * He, nx + nw * He))
    i = 0
    for j in range(1, He + 1):
        rows = slice(i * nx, nx * He)
        EX[rows, j * nw:(j + 1) * nw] = Apow_v[:nx * He - i * nx]
        i += 1
    EX[:, :nx] = Apow_v[nx:]
    nCAB = np.vstack([np.zeros((nym, nu))] + [nCA[i] @ Bu for i in range(He)])
    G = np.zeros((nym * He, nu * He))
    col_begin, col_end = (1, He - 1) if p == 0 else (0, He - 2)
    i = 0
    for j in range(col_begin, col_end + 1):
        rows = slice(i * nym, nym * He)
        G[rows, j * nu:(j + 1) * nu] = nCAB[:nym * He - i * nym]
        i += 1
    if p == 0:
        G[:, :nu] = nCAB[nym:]
    AB = np.vstack([Apow[i] @ Bu for i in range(He)])
    GX = np.zeros((nx * He, nu * He))
    for j in range(He):
        GX[j * nx:, j * nu:(j + 1) * nu] = AB[:nx * He - j * nx]
    nCABd = np.vstack([-Ddm] + [nCA[i] @ Bd for i in range(He)])
    J = np.zeros((nym * He, nd * (He + 1)))
    i = 0
    for j in range(1, He + 1):
        rows = slice(i * nym, nym * He)
        J[rows, j * nd:(j + 1) * nd] = nCABd[:nym * He - i * nym]
        i += 1
    if p == 0:
        J[:, :nd] = nCABd[nym:]
    ABd = np.vstack([Apow[i] @ Bd for i in range(He)])
    JX = np.zeros((nx * He, nd * (He + 1)))
    for j in range(He):
        JX[j * nx:, (j + p) * nd:(j + p + 1) * nd] = ABd[:nx * He - j * nx]
    Acs = np.cumsum(np.array(Apow), axis=0)
    coefB = np.zeros((nym * He, nx))
    row_begin, row_end = (0, He - 1) if p == 0 else (1, He - 2)
    j = 0
    for i in range(row_begin, row_end + 1):
        coefB[i * nym:(i + 1) * nym] = -Cm @ Acs[j]
        j += 1
    B = coefB @ f_minus_x
    BX = np.vstack([Acs[j] for j in range(He)]) @ f_minus_x
    return E, G, J, B, exbar, EX, GX, JX, BX


class MovingHorizonEstimator(StateEstimator):
    """Linear MHE (src/estimator/mhe/construct.jl:74-241, kw constructor :528-630; default arrival
    covariance estimator = KalmanFilter, :642-648)."""

    def __init__(self, model, He, i_ym=None, sigmaP_0=None, sigmaQ=None, sigmaR=None, nint_u=0, nint_ym=None,
                 sigmaPint_u_0=None, sigmaQint_u=None, sigmaPint_ym_0=None, sigmaQint_ym=None, Cwt=np.inf,
                 direct=True, P0hat=None, Qhat=None, Rhat=None):
        self._init_common(model, i_ym, nint_u, nint_ym)
        self.He, self.direct = int(He), bool(direct)
        if self.He < 1:
            raise ValueError("Estimation horizon He should be >= 1")
        nx, nxh, nym = model.nx, self.nxhat, len(self.i_ym)
        self.nym = nym
        one = lambda v, n, d: np.full(n, d) if v is None else np.asarray(v, float).reshape(n)
        nsu, nsy = int(np.sum(self.nint_u)), int(np.sum(self.nint_ym))
        sP = np.concatenate([one(sigmaP_0, nx, 1 / nx), one(sigmaPint_u_0, nsu, 1.0), one(sigmaPint_ym_0, nsy, 1.0)])
        sQ = np.concatenate([one(sigmaQ, nx, 1 / nx), one(sigmaQint_u, nsu, 1.0), one(sigmaQint_ym, nsy, 1.0)])
        sR = one(sigmaR, nym, 1.0)
        self.P0hat, self.Qhat, self.Rhat = np.diag(sP ** 2), np.diag(sQ ** 2), np.diag(sR ** 2)
        # full covariance matrices (the reference's second constructor, src/estimator/mhe/construct.jl:632-660)
        if P0hat is not None: self.P0hat = np.atleast_2d(np.asarray(P0hat, float))
        if Qhat is not None: self.Qhat = np.atleast_2d(np.asarray(Qhat, float))
        if Rhat is not None: self.Rhat = np.atleast_2d(np.asarray(Rhat, float))
        self.Cwt = float(Cwt)
        if self.Cwt < 0:
            raise ValueError("Cwt weight should be >= 0")  # mhe/construct.jl (ArgumentError)
        self.neps = 0 if np.isinf(self.Cwt) else 1
        f = self.fophat - self.xophat
        (self.E, self.G, self.J, self.B, self.exbar, self.EX, self.GX, self.JX, self.BX) = init_predmat_mhe(
            self.He, self.Ahat, self.Buhat, self.Cmhat, self.Bdhat, self.Ddmhat, f, self.direct)
        He, nw, nu, nd = self.He, nxh, model.nu, model.nd
        inf = np.inf
        self.con = dict(xhat0min=np.full(nxh, -inf), xhat0max=np.full(nxh, inf),
                        X0min=np.full(nxh * He, -inf), X0max=np.full(nxh * He, inf),
                        Wmin=np.full(nw * He, -inf), Wmax=np.full(nw * He, inf),
                        Vmin=np.full(nym * He, -inf), Vmax=np.full(nym * He, inf),
                        c_xmin=np.zeros(nxh), c_xmax=np.zeros(nxh), C_xmin=np.zeros(nxh * He),
                        C_xmax=np.zeros(nxh * He), C_wmin=np.zeros(nw * He), C_wmax=np.zeros(nw * He),
                        C_vmin=np.zeros(nym * He), C_vmax=np.zeros(nym * He))
        self.nZ = self.neps + nxh + nw * He
        self.invQ_He = np.kron(np.eye(He), np.linalg.inv(self.Qhat))
        self.invR_He = np.kron(np.eye(He), np.linalg.inv(self.Rhat))
        self.solved_once = False
        self.reset()

    def reset(self):
        """init_estimate_cov! (src/estimator/mhe/execute.jl:2-37) with zero inputs."""
        He, nxh, nu, nd, nym = self.He, self.nxhat, self.model.nu, self.model.nd, self.nym
        self.Ztilde = np.zeros(self.nZ)
        self.Y0m = np.full(nym * He, np.nan)
        self.U0 = np.full(nu * He, np.nan)
        self.D0 = np.full(nd * (He + 1), np.nan)
        self.X0_old = np.full(nxh * He, np.nan)
        if self.direct:
            self.U0[:nu] = 0.0
        if nd:
            self.D0[:nd] = 0.0
        self.lastu0 = np.zeros(nu)
        self.xhat0 = np.zeros(nxh)
        self.xhat0arr_old = np.zeros(nxh)
        self.Parr_old = self.P0hat.copy()
        self.invPbar = np.linalg.inv(self.Parr_old)
        self.Nk = 0

    def setconstraint(self, xhatmin=None, xhatmax=None, whatmin=None, whatmax=None, vhatmin=None, vhatmax=None,
                      c_xhatmin=None, c_xhatmax=None, c_whatmin=None, c_whatmax=None, c_vhatmin=None,
                      c_vhatmax=None, Xhatmin=None, Xhatmax=None, Whatmin=None, Whatmax=None, Vhatmin=None, Vhatmax=None,
                      C_xhatmin=None, C_xhatmax=None, C_whatmin=None, C_whatmax=None, C_vhatmin=None, C_vhatmax=None):
        """src/estimator/mhe/construct.jl:858-1046.  Lower-case keywords hold for every stage of the window, capitalised
        ones give the whole window: X̂ over He+1 stages (the first one is the arrival state x̂0(k-Nk+p)), Ŵ and V̂ over He."""
        c, He, nxh, nym = self.con, self.He, self.nxhat, self.nym

        def a(v, n, name):
            v = np.asarray(v, float).reshape(-1)
            if v.size != n:
                raise ValueError(f"{name} size must be ({n},)")  # DimensionMismatch
            return v
        Xop = np.tile(self.xophat, He)
        pattern = lambda: tuple(np.isfinite(c[k]).tobytes() for k in ("xhat0min", "xhat0max", "X0min", "X0max", "Wmin", "Wmax",
                                                                      "Vmin", "Vmax"))
        old_pattern, old = pattern(), {k: np.array(v, copy=True) for k, v in c.items()}
        try:
            if Xhatmin is not None:
                v = a(Xhatmin, nxh * (He + 1), "Xhatmin")
                c["xhat0min"], c["X0min"] = v[:nxh] - self.xophat, v[nxh:] - Xop
            elif xhatmin is not None:
                v = a(xhatmin, nxh, "xhatmin")
                c["xhat0min"], c["X0min"] = v - self.xophat, np.tile(v, He) - Xop
            if Xhatmax is not None:
                v = a(Xhatmax, nxh * (He + 1), "Xhatmax")
                c["xhat0max"], c["X0max"] = v[:nxh] - self.xophat, v[nxh:] - Xop
            elif xhatmax is not None:
                v = a(xhatmax, nxh, "xhatmax")
                c["xhat0max"], c["X0max"] = v - self.xophat, np.tile(v, He) - Xop
            for key, small, big, n in (("Wmin", whatmin, Whatmin, nxh), ("Wmax", whatmax, Whatmax, nxh),
                                       ("Vmin", vhatmin, Vhatmin, nym), ("Vmax", vhatmax, Vhatmax, nym)):
                if big is not None:
                    c[key] = a(big, n * He, key).copy()
                elif small is not None:
                    c[key] = np.tile(a(small, n, key), He)
            ecr = [c_xhatmin, c_xhatmax, c_whatmin, c_whatmax, c_vhatmin, c_vhatmax,
                   C_xhatmin, C_xhatmax, C_whatmin, C_whatmax, C_vhatmin, C_vhatmax]
            if any(e is not None for e in ecr):
                if not self.neps:
                    raise ValueError("Slack variable weight Cwt must be finite to set softness parameters")
                if self.solved_once:
                    raise RuntimeError("Cannot set softness parameters after calling updatestate!")

                def soft(small, big, n, reps, name):
                    if big is not None:
                        v = a(big, n * reps, "C_" + name)
                    elif small is not None:
                        v = np.tile(a(small, n, "c_" + name), reps)
                    else:
                        return None
                    if (v < 0).any():
                        raise ValueError(f"C_{name} weights should be non-negative")
                    return v
                for lo, hi, small, big, name in (("c_xmin", "C_xmin", c_xhatmin, C_xhatmin, "xhatmin"),
                                                 ("c_xmax", "C_xmax", c_xhatmax, C_xhatmax, "xhatmax")):
                    v = soft(small, big, nxh, He + 1, name)
                    if v is not None:
                        c[lo], c[hi] = v[:nxh].copy(), v[nxh:].copy()
                for key, small, big, n, name in (("C_wmin", c_whatmin, C_whatmin, nxh, "whatmin"),
                                                 ("C_wmax", c_whatmax, C_whatmax, nxh, "whatmax"),
                                                 ("C_vmin", c_vhatmin, C_vhatmin, nym, "vhatmin"),
                                                 ("C_vmax", c_vhatmax, C_vhatmax, nym, "vhatmax")):
                    v = soft(small, big, n, He, name)
                    if v is not None:
                        c[key] = v
            if self.solved_once and pattern() != old_pattern:  # construct.jl:1037-1039
                raise RuntimeError("Cannot modify +-Inf constraints after first solve of estimation problem")
        except Exception:
            c.update(old)  # nothing is stored by a call that fails
            raise
        return self

    # ---- windows (add_data_windows!, execute.jl:497-547) ----
    def add_data_windows(self, y0m, d0, u0):
        nxh, nym, nd, nu, He = self.nxhat, self.nym, self.model.nd, self.model.nu, self.He
        x_old = self.xhat0.copy()
        self.Nk += 1
        Nk = self.Nk
        ismoving = Nk > He
        if ismoving:
            self.Y0m[:-nym] = self.Y0m[nym:].copy()
            self.Y0m[-nym:] = y0m
            if nd:
                self.D0[:-nd] = self.D0[nd:].copy()
                self.D0[-nd:] = d0
            self.U0[:-nu] = self.U0[nu:].copy()
            self.U0[-nu:] = u0
            self.X0_old[:-nxh] = self.X0_old[nxh:].copy()
            self.X0_old[-nxh:] = x_old
            self.Nk = He
        else:
            self.Y0m[nym * (Nk - 1):nym * Nk] = y0m
            if nd:
                self.D0[nd * Nk:nd * (Nk + 1)] = d0
            self.U0[nu * (Nk - 1):nu * Nk] = u0
            self.X0_old[nxh * (Nk - 1):nxh * Nk] = x_old
        self.xhat0arr_old = self.X0_old[:nxh].copy()
        return ismoving

    # ---- arrival covariance (correct_cov!/update_cov!/invert_cov!, execute.jl:729-797) ----
    def _kf_correct(self):
        P, Cm = self.Parr_old, self.Cmhat
        M = Cm @ P @ Cm.T + self.Rhat
        K = np.linalg.solve(M.T, (P @ Cm.T).T).T
        return (np.eye(self.nxhat) - K @ Cm) @ P

    def _set_cov(self, Pnew):
        if not np.all(np.isfinite(Pnew)):
            return  # keeps the old one (execute.jl:737-750)
        Pnew = np.tril(Pnew) + np.tril(Pnew, -1).T  # Hermitian(:L)
        try:
            np.linalg.cholesky(Pnew)
        except np.linalg.LinAlgError:
            self.Parr_old = Pnew
            return
        self.Parr_old = Pnew
        self.invPbar = np.linalg.inv(Pnew)

    def correct_cov(self):
        self._set_cov(self._kf_correct())

    def update_cov(self):
        P = self.Parr_old if self.direct else self._kf_correct()  # KalmanFilter.update_estimate!, kalman.jl:520-525
        self._set_cov(self.Ahat @ P @ self.Ahat.T + self.Qhat)

    # ---- the QP of the current window (initpred! :419-457, linconstraint! transcription.jl:732-781) ----
    def build_qp(self):
        Nk, He, nxh, nym, neps, nu, nd = self.Nk, self.He, self.nxhat, self.nym, self.neps, self.model.nu, self.model.nd
        nw = nxh
        nZ = nxh + nw * Nk
        U0, Y0m, D0 = self.U0[:nu * Nk], self.Y0m[:nym * Nk], self.D0[:nd * (Nk + 1)]
        E = self.E[:nym * Nk, :nZ].copy()
        F = Y0m + self.B[:nym * Nk] + self.G[:nym * Nk, :nu * Nk] @ U0
        if nd:
            F = F + self.J[:nym * Nk, :nd * (Nk + 1)] @ D0
        nan = np.isnan(F)
        E[nan] = 0.0
        F = np.where(nan, 0.0, F)
        fxbar = self.xhat0arr_old
        z = lambda r, c: np.zeros((r, c))
        Et = np.hstack([z(nym * Nk, neps), E])
        ext = np.hstack([z(nxh, neps), self.exbar[:, :nZ]])
        EZ, FZ = np.vstack([ext, Et]), np.concatenate([fxbar, F])
        M = np.block([[self.invPbar, z(nxh, nym * Nk)], [z(nym * Nk, nxh), self.invR_He[:nym * Nk, :nym * Nk]]])
        Tw = np.hstack([z(nw * Nk, nxh), np.eye(nw * Nk)])
        Nt = np.zeros((neps + nZ, neps + nZ))
        if neps:
            Nt[0, 0] = self.Cwt
        Nt[neps:, neps:] = Tw.T @ self.invQ_He[:nw * Nk, :nw * Nk] @ Tw
        H = 2 * (EZ.T @ M @ EZ + Nt)
        q = 2 * (M @ EZ).T @ FZ
        r = FZ @ M @ FZ
        FX = self.BX[:nxh * Nk] + self.GX[:nxh * Nk, :nu * Nk] @ U0
        if nd:
            FX = FX + self.JX[:nxh * Nk, :nd * (Nk + 1)] @ D0
        EXt = np.hstack([z(nxh * Nk, neps), self.EX[:nxh * Nk, :nZ]])
        c = self.con
        col = lambda v: np.asarray(v, float).reshape(-1, 1)
        tr = lambda b, n: b[-n * Nk:] if Nk < He else b     # trunc_bounds (execute.jl:550-564)
        trc = lambda b, n: b[:n * Nk]
        ex = -self.exbar[:, :nZ]                                # relaxarrival: ex̂ = -ex̄
        rows, rhs = [], []

        def add(Amat, cvec, bvec, sign):
            cvec = np.zeros(len(bvec)) if not neps else cvec
            Afull = np.hstack([-col(cvec), sign * Amat]) if neps else sign * Amat
            fin = np.isfinite(bvec)
            rows.append(Afull[fin])
            rhs.append(bvec[fin])
        add(ex, c["c_xmin"], -c["xhat0min"], -1.0)
        add(ex, c["c_xmax"], c["xhat0max"], +1.0)
        add(self.EX[:nxh * Nk, :nZ], trc(c["C_xmin"], nxh), -tr(c["X0min"], nxh) + FX, -1.0)
        add(self.EX[:nxh * Nk, :nZ], trc(c["C_xmax"], nxh), tr(c["X0max"], nxh) - FX, +1.0)
        add(Tw, trc(c["C_wmin"], nw), -tr(c["Wmin"], nw), -1.0)
        add(Tw, trc(c["C_wmax"], nw), tr(c["Wmax"], nw), +1.0)
        add(E, trc(c["C_vmin"], nym), -tr(c["Vmin"], nym) + F, -1.0)
        add(E, trc(c["C_vmax"], nym), tr(c["Vmax"], nym) - F, +1.0)
        A = np.vstack(rows) if rows else np.zeros((0, neps + nZ))
        b = np.concatenate(rhs) if rhs else np.zeros(0)
        lb = np.full(neps + nZ, -np.inf)
        if neps:
            lb[0] = 0.0
        return dict(H=H, q=q, r=r, A=A, b=b, lb=lb, F=F, FX=FX, EXt=EXt, Et=Et, nZ=nZ, EZ=EZ, FZ=FZ, Mhat=M, Nt=Nt)

    def solve_window(self):
        """initpred! + linconstraint! + optim_objective! + getstate! (execute.jl:44-55, 576-638)."""
        P = self.build_qp()
        neps, nxh, Nk = self.neps, self.nxhat, self.Nk
        # warm start (fallback on error), set_warmstart_mhe!, transcription.jl:967-1001
        nxt, nW = neps + nxh, nxh * self.He
        Zs = np.zeros(self.nZ)
        if neps:
            Zs[0] = self.Ztilde[0]
        Zs[neps:nxt] = self.xhat0arr_old
        Zs[nxt:nxt + nW - nxh] = self.Ztilde[nxt + nxh:nxt + nW]
        sol = _qp.solve_qp(P["H"], P["q"], P["A"], P["b"], P["lb"], None)
        self.last_qp, self.solved_once = sol, True
        Z = np.zeros(self.nZ)
        if sol["status"] == _qp.INFEASIBLE:
            Z[:] = Zs
        else:
            Z[:neps + P["nZ"]] = sol["z"]
        Z[nxt + nxh * Nk:] = 0.0  # fill0unused!
        self.Ztilde = Z
        X0 = P["EXt"] @ Z[:neps + P["nZ"]] + P["FX"]
        self.Vhat = P["Et"] @ Z[:neps + P["nZ"]] + P["F"]
        self.X0 = X0
        self.xhat0 = X0[(Nk - 1) * nxh:Nk * nxh].copy()
        self.Jval = 0.5 * Z[:neps + P["nZ"]] @ P["H"] @ Z[:neps + P["nZ"]] + P["q"] @ Z[:neps + P["nZ"]] + P["r"]

    def correct_estimate(self, y0m, d0):
        if self.direct:
            ismoving = self.add_data_windows(y0m, d0, self.lastu0)
            if ismoving:
                self.correct_cov()
            self.solve_window()

    def update_estimate(self, u0, y0m, d0):
        if not self.direct:
            self.add_data_windows(y0m, d0, u0)
            self.solve_window()
        if self.Nk == self.He:
            self.update_cov()
        self.lastu0 = np.asarray(u0, float).copy()


class KalmanFilter(StateEstimator):
    """Time-varying Kalman filter (src/estimator/kalman.jl:311-525, 1235-1290): the reference against
    which the unconstrained linear MHE is asserted equal (test/2_test_state_estim.jl:1750-1784)."""

    def __init__(self, model, i_ym=None, sigmaP_0=None, sigmaQ=None, sigmaR=None, nint_u=0, nint_ym=None,
                 sigmaPint_u_0=None, sigmaQint_u=None, sigmaPint_ym_0=None, sigmaQint_ym=None, direct=True):
        self._init_common(model, i_ym, nint_u, nint_ym)
        self.direct = direct
        nx, nym = model.nx, len(self.i_ym)
        one = lambda v, n, d: np.full(n, d) if v is None else np.asarray(v, float).reshape(n)
        nsu, nsy = int(np.sum(self.nint_u)), int(np.sum(self.nint_ym))
        sP = np.concatenate([one(sigmaP_0, nx, 1 / nx), one(sigmaPint_u_0, nsu, 1.0), one(sigmaPint_ym_0, nsy, 1.0)])
        sQ = np.concatenate([one(sigmaQ, nx, 1 / nx), one(sigmaQint_u, nsu, 1.0), one(sigmaQint_ym, nsy, 1.0)])
        self.Phat, self.Qhat, self.Rhat = np.diag(sP ** 2), np.diag(sQ ** 2), np.diag(one(sigmaR, nym, 1.0) ** 2)

    def setstate(self, xhat, Phat=None):
        """setstate!(estim, x̂, P̂) (src/estimator/execute.jl:424-438)."""
        xhat = np.asarray(xhat, float).reshape(-1)
        if xhat.size != self.nxhat:
            raise ValueError(f"xhat size must be ({self.nxhat},)")
        self.xhat0 = xhat - self.xophat
        if Phat is not None:
            Phat = np.asarray(Phat, float)
            if Phat.shape != (self.nxhat, self.nxhat):
                raise ValueError(f"Phat size must be ({self.nxhat}, {self.nxhat})")
            self.Phat = np.tril(Phat) + np.tril(Phat, -1).T

    def correct_estimate(self, y0m, d0):
        if np.isnan(y0m).any():  # kalman.jl:478-482: the correction is skipped
            return
        P, Cm = self.Phat, self.Cmhat
        M = Cm @ P @ Cm.T + self.Rhat
        K = np.linalg.solve(M.T, (P @ Cm.T).T).T
        self.xhat0 = self.xhat0 + K @ (y0m - (Cm @ self.xhat0 + self.Ddmhat @ d0))
        self.Khat = K
        Pc = (np.eye(self.nxhat) - K @ Cm) @ P
        self.Phat = np.tril(Pc) + np.tril(Pc, -1).T           # Hermitian(P̂corr, :L), kalman.jl:1266

    def update_estimate(self, u0, y0m, d0):
        if not self.direct:
            self.correct_estimate(y0m, d0)
        self.xhat0 = self.Ahat @ self.xhat0 + self.Buhat @ u0 + self.Bdhat @ d0 + self.fophat - self.xophat
        Pn = self.Ahat @ self.Phat @ self.Ahat.T + self.Qhat
        self.Phat = np.tril(Pn) + np.tril(Pn, -1).T           # Hermitian(P̂next, :L), kalman.jl:1288
