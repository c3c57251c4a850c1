"""numpy fp64 restatement of the reference's linear ``MovingHorizonEstimator`` with the ``MultipleShooting``
transcription (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Decision vector ``Z̃ = [ε; x̂0(k-Nk+p); X̂0; Ŵ]`` (``get_nZ_mhe`` src/estimator/mhe/transcription.jl:3): the stage states
are decision variables and the model enters as the linear EQUALITY constraints ``ES Z + FS = 0`` (``init_defectmat_mhe``
:465-488, ``linconstrainteq!`` :852-883).  The equality-constrained QP is solved EXACTLY by ``oracle.qp.solve_qp_eq``
(generic null-space elimination: nothing of the single-shooting condensation is used), which makes this class an
independent check of the condensed estimator the CUDA path runs.  Pinned to the reference's known answers
test/2_test_state_estim.jl:1126-1139 and :1722-1733 (tests/test_oracle_mhe_ms.py).
"""
from __future__ import annotations

import numpy as np

from . import qp as _qp
from .mhe import MovingHorizonEstimator


def _repeatdiag(M, n):
    return np.kron(np.eye(n), M)


def init_predmat_mhe_ms(He, A, Cm, Ddm, nu, nd, direct):
    """transcription.jl:327-357 (LinModel + MultipleShooting): E, G, J, B, exbar, EX, GX, JX, BX."""
    nym, nx = Cm.shape
    nw, p = nx, (0 if direct else 1)
    nX, nW, nV, nU, nD = nx * He, nw * He, nym * He, nu * He, nd * (He + 1)
    z = np.zeros
    E = np.hstack([z((nV, (1 - p) * nx)), _repeatdiag(-Cm, He), z((nV, p * nx + nW))])
    exbar = np.hstack([-np.eye(nx), z((nx, nX + nW))])
    EX = np.hstack([z((nX, nx)), np.eye(nX), z((nX, nW))])
    G, GX = z((nV, nU)), z((nX, nU))
    J = np.hstack([z((nV, nd)), _repeatdiag(-Ddm, He)]) if nd else z((nV, 0))
    JX = z((nX, nD))
    return E, G, J, z(nV), exbar, EX, GX, JX, z(nX)


def init_defectmat_mhe(He, A, Bu, Bd, f_minus_x, direct):
    """transcription.jl:465-488: ES, GS, JS, BS of  Ŝ = ES Z + GS U0 + JS D0 + BS  (= 0)."""
    nx, nd = A.shape[0], Bd.shape[1]
    nw, nX, p = nx, nx * He, (0 if direct else 1)
    ES = np.hstack([np.zeros((nX, nx)), _repeatdiag(-np.eye(nx), He), _repeatdiag(np.eye(nw), He)])
    for j in range(He):
        ES[j * nx:(j + 1) * nx, j * nx:(j + 1) * nx] = A
    GS = _repeatdiag(Bu, He)
    JS = np.hstack([np.zeros((nX, p * nd)), _repeatdiag(Bd, He), np.zeros((nX, (1 - p) * nd))])
    BS = np.tile(f_minus_x, He)
    return ES, GS, JS, BS


class MovingHorizonEstimatorMS(MovingHorizonEstimator):
    """``MovingHorizonEstimator(model; He, transcription=MultipleShooting())`` for a LinModel.  Windows, arrival covariance
    and the per-period call sequence are inherited (they do not depend on the transcription, mhe/execute.jl:44-84)."""

    def __init__(self, model, He, **kw):
        super().__init__(model, He, **kw)
        nu, nd = model.nu, model.nd
        f = self.fophat - self.xophat
        (self.E, self.G, self.J, self.B, self.exbar, self.EX, self.GX, self.JX, self.BX) = init_predmat_mhe_ms(
            self.He, self.Ahat, self.Cmhat, self.Ddmhat, nu, nd, self.direct)
        self.ES, self.GS, self.JS, self.BS = init_defectmat_mhe(self.He, self.Ahat, self.Buhat, self.Bdhat, f, self.direct)
        self.nZ = self.neps + self.nxhat + 2 * self.nxhat * self.He
        self.Ztilde = np.zeros(self.nZ)

    def reset(self):
        super().reset()
        if hasattr(self, "ES"):
            self.Ztilde = np.zeros(self.nZ)

    def _i_Z_Nk(self):
        """get_i_Z̃_Nk (transcription.jl:6-13) without the slack: columns of the window's variables."""
        nxh, He, Nk = self.nxhat, self.He, self.Nk
        return np.r_[0:nxh + nxh * Nk, nxh + nxh * He:nxh + nxh * He + nxh * Nk]

    def build_qp(self):
        Nk, He, nxh, nym, neps, nu, nd = self.Nk, self.He, self.nxhat, self.nym, self.neps, self.model.nu, self.model.nd
        nw = nxh
        iz = self._i_Z_Nk()
        nZ = iz.size
        U0, Y0m, D0 = self.U0[:nu * Nk], self.Y0m[:nym * Nk], self.D0[:nd * (Nk + 1)]
        E = self.E[:nym * Nk][:, iz].copy()
        F = Y0m + self.B[:nym * Nk] + self.G[:nym * Nk, :nu * Nk] @ U0
        if nd:
            F = F + self.J[:nym * Nk, :nd * (Nk + 1)] @ D0
        nan = np.isnan(F)
        E[nan] = 0.0
        F = np.where(nan, 0.0, F)
        z = lambda r, c: np.zeros((r, c))
        exb = self.exbar[:, iz]
        Et = np.hstack([z(nym * Nk, neps), E])
        ext = np.hstack([z(nxh, neps), exb])
        EZ, FZ = np.vstack([ext, Et]), np.concatenate([self.xhat0arr_old, F])
        M = np.block([[self.invPbar, z(nxh, nym * Nk)], [z(nym * Nk, nxh), self.invR_He[:nym * Nk, :nym * Nk]]])
        Tw = np.hstack([z(nw * Nk, nxh + nxh * Nk), np.eye(nw * Nk)])
        Nt = np.zeros((neps + nZ, neps + nZ))
        if neps:
            Nt[0, 0] = self.Cwt
        Nt[neps:, neps:] = Tw.T @ self.invQ_He[:nw * Nk, :nw * Nk] @ Tw
        H = 2 * (EZ.T @ M @ EZ + Nt)
        q = 2 * (M @ EZ).T @ FZ
        r = FZ @ M @ FZ
        # defects (trunc_defectmat, execute.jl:684-704; linconstrainteq!, transcription.jl:852-883)
        ES = self.ES[:nxh * Nk][:, iz]
        FS = self.BS[:nxh * Nk] + self.GS[:nxh * Nk, :nu * Nk] @ U0
        if nd:
            FS = FS + self.JS[:nxh * Nk, :nd * (Nk + 1)] @ D0
        Aeq = np.hstack([z(nxh * Nk, neps), ES])
        EXw = self.EX[:nxh * Nk][:, iz]
        c = self.con
        col = lambda v: np.asarray(v, float).reshape(-1, 1)
        tr = lambda b, n: b[-n * Nk:] if Nk < He else b     # trunc_bounds (execute.jl:550-564)
        trc = lambda b, n: b[:n * Nk]
        rows, rhs = [], []

        def add(Amat, cvec, bvec, sign):
            cvec = np.zeros(len(bvec)) if not neps else cvec
            Afull = np.hstack([-col(cvec), sign * Amat]) if neps else sign * Amat
            fin = np.isfinite(bvec)
            rows.append(Afull[fin])
            rhs.append(bvec[fin])
        # (state and noise bounds are box constraints on Z̃ in the reference when hard, boxconstraint_states!
        # transcription.jl:670-687 -- the same feasible set as these rows)
        add(-exb, c["c_xmin"], -c["xhat0min"], -1.0)
        add(-exb, c["c_xmax"], c["xhat0max"], +1.0)
        add(EXw, trc(c["C_xmin"], nxh), -tr(c["X0min"], nxh), -1.0)
        add(EXw, trc(c["C_xmax"], nxh), tr(c["X0max"], nxh), +1.0)
        add(Tw, trc(c["C_wmin"], nw), -tr(c["Wmin"], nw), -1.0)
        add(Tw, trc(c["C_wmax"], nw), tr(c["Wmax"], nw), +1.0)
        add(E, trc(c["C_vmin"], nym), -tr(c["Vmin"], nym) + F, -1.0)
        add(E, trc(c["C_vmax"], nym), tr(c["Vmax"], nym) - F, +1.0)
        A = np.vstack(rows) if rows else np.zeros((0, neps + nZ))
        b = np.concatenate(rhs) if rhs else np.zeros(0)
        lb = np.full(neps + nZ, -np.inf)
        if neps:
            lb[0] = 0.0
        return dict(H=H, q=q, r=r, A=A, b=b, lb=lb, F=F, Et=Et, nZ=nZ, Aeq=Aeq, beq=-FS, iz=iz, EXw=EXw)

    def solve_window(self):
        P = self.build_qp()
        neps, nxh, Nk, He = self.neps, self.nxhat, self.Nk, self.He
        sol = _qp.solve_qp_eq(P["H"], P["q"], P["A"], P["b"], P["Aeq"], P["beq"], P["lb"], None)
        self.last_qp, self.solved_once = sol, True
        idx = np.r_[0:neps, neps + P["iz"]]
        Z = np.zeros(self.nZ)
        if sol["status"] == _qp.INFEASIBLE or not np.all(np.isfinite(sol["z"])):
            # set_warmstart_mhe! (transcription.jl:1037-1076): shifted previous solution
            nxt, nX, nW = neps + nxh, nxh * He, nxh * He
            Zs = np.zeros(self.nZ)
            if neps:
                Zs[0] = self.Ztilde[0]
            Zs[neps:nxt] = self.xhat0arr_old
            Zs[nxt:nxt + nX - nxh] = self.Ztilde[nxt + nxh:nxt + nX]
            Zs[nxt + nX - nxh:nxt + nX] = self.Ztilde[nxt + nX - nxh:nxt + nX]
            Zs[nxt + nX:nxt + nX + nW - nxh] = self.Ztilde[nxt + nX + nxh:nxt + nX + nW]
            Z[:] = Zs
        else:
            Z[idx] = sol["z"]
        nxt = neps + nxh
        Z[nxt + nxh * Nk:nxt + nxh * He] = 0.0          # fill0unused! (transcription.jl:1084-1090)
        Z[nxt + nxh * He + nxh * Nk:] = 0.0
        self.Ztilde = Z
        zw = Z[idx]
        X0 = P["EXw"] @ zw[neps:]
        self.Vhat = P["Et"] @ zw + P["F"]
        self.X0 = X0
        self.xhat0 = X0[(Nk - 1) * nxh:Nk * nxh].copy()
        self.Jval = 0.5 * zw @ P["H"] @ zw + P["q"] @ zw + P["r"]
