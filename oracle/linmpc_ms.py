"""numpy fp64 restatement of ``LinMPC`` with the ``MultipleShooting`` transcription (TEST INFRASTRUCTURE ONLY).

SURVEY.md section 8(f-3).  Decision vector Z = [ΔU; X̂0] (src/controller/transcription.jl:28-60), Z̃ = [Z; ε]:
  * init_ZtoΔU / init_ZtoU with nZ = nu Hc + nx̂ Hp columns (construct.jl:743-809),
  * init_predmat for LinModel + MultipleShooting (transcription.jl:217-240): E = [0 diag(Ĉ)], J = diag(D̂d), ex̂ = [0 I],
    every other matrix zero,
  * init_defectmat (transcription.jl:373-414): ES, GS, JS, KS, VS, BS of the equality constraints ES Z + FS = 0,
  * linconstrainteq! (transcription.jl:913-928): FS = GS d0 + JS D̂0 + KS x̂0 + VS u0(k-1) + BS, beq = -FS,
  * box constraints: hard ΔU bounds and hard terminal bounds become bounds on Z̃ (boxconstraint_terminal!, :755-775) and
    leave i_b (deleteΔU_lincon!, deletex̂end_lincon!, :777-797),
  * warm start for "other transcriptions" (transcription.jl:1089-1102),
  * the objective, initpred!, linconstraint!, getinput!, getinfo of the common path (execute.jl).
The QP is the reference's actual MultipleShooting QP (equality constrained, semidefinite Hessian); it is solved
exactly by ``oracle.qp.solve_qp_eq`` (generic null-space elimination).  Pinned to test/3_test_predictive_control.jl:120-127
and :570-579.  Diagonal or dense weights; no custom linear constraints (the mirror does not take them with MS either).
"""
from __future__ import annotations

import numpy as np

from . import qp as _qp
from .linmpc import (DEFAULT_CWT, DEFAULT_HC, DEFAULT_HP0, DEFAULT_LWT, DEFAULT_MWT, DEFAULT_NWT, StateEstimator,
                     SteadyKalmanFilter, init_ZtoU, move_blocking)


def init_predmat_ms(estim, Hp, Hc):
    """src/controller/transcription.jl:217-240."""
    nu, nx, ny, nd = estim.model.nu, estim.nxhat, estim.model.ny, estim.model.nd
    E = np.hstack([np.zeros((Hp * ny, Hc * nu)), np.kron(np.eye(Hp), estim.Chat)])
    ex = np.hstack([np.zeros((nx, Hc * nu + (Hp - 1) * nx)), np.eye(nx)])
    J = np.kron(np.eye(Hp), estim.Ddhat) if nd else np.zeros((Hp * ny, 0))
    return E, J, ex


def init_defectmat(estim, Hp, Hc, nb):
    """src/controller/transcription.jl:373-414."""
    nu, nx, nd = estim.model.nu, estim.nxhat, estim.model.nd
    A, Bu, Bd = estim.Ahat, estim.Buhat, estim.Bdhat
    KS = np.vstack([A, np.zeros((nx * (Hp - 1), nx))])
    VS = np.tile(Bu, (Hp, 1))
    ES = np.hstack([np.zeros((nx * Hp, nu * Hc)), -np.eye(nx * Hp)])
    for j in range(Hc):
        for i in range(j, Hc):
            r0 = nx * int(np.sum(nb[:i]))
            for l in range(nb[i]):
                ES[r0 + nx * l:r0 + nx * (l + 1), nu * j:nu * (j + 1)] = Bu
    for j in range(1, Hp):
        ES[nx * j:nx * (j + 1), nu * Hc + nx * (j - 1):nu * Hc + nx * j] = A
    GS = np.vstack([Bd, np.zeros((nx * (Hp - 1), nd))])
    JS = np.zeros((nx * Hp, nd * Hp))
    for j in range(1, Hp):
        JS[nx * j:nx * (j + 1), nd * (j - 1):nd * j] = Bd
    BS = np.tile(estim.fophat - estim.xophat, Hp)
    return ES, GS, JS, KS, VS, BS


class LinMPCMultipleShooting:
    def __init__(self, model_or_estim, Hp=None, Hc=DEFAULT_HC, Mwt=None, Nwt=None, Lwt=None, Cwt=DEFAULT_CWT, **kwargs):
        estim = model_or_estim if isinstance(model_or_estim, StateEstimator) else SteadyKalmanFilter(model_or_estim, **kwargs)
        model = estim.model
        self.estim, self.model = estim, model
        nu, ny, nd, nx = model.nu, model.ny, model.nd, estim.nxhat
        if Hp is None:
            Hp = DEFAULT_HP0 + int(np.sum(np.isclose(np.abs(np.linalg.eigvals(model.A)), 0.0, atol=1e-3)))
        nb = move_blocking(Hp, Hc)
        Hc = len(nb)
        self.Hp, self.Hc, self.nb = Hp, Hc, nb
        w = lambda v, d, n: np.full(n, d) if v is None else np.asarray(v, float).reshape(n)
        self.Mwt, self.Nwt, self.Lwt = w(Mwt, DEFAULT_MWT, ny), w(Nwt, DEFAULT_NWT, nu), w(Lwt, DEFAULT_LWT, nu)
        self.Cwt = float(Cwt)
        self.neps = neps = 0 if np.isinf(self.Cwt) else 1
        self.nDU, self.nX = nu * Hc, nx * Hp
        self.nZ = nZ = self.nDU + self.nX
        self.n = nZ + neps
        inf = np.inf
        self.con = dict(U0min=np.full(nu * Hp, -inf), U0max=np.full(nu * Hp, inf), DUmin=np.full(self.nDU, -inf),
                        DUmax=np.full(self.nDU, inf), Y0min=np.full(ny * Hp, -inf), Y0max=np.full(ny * Hp, inf),
                        xhat0min=np.full(nx, -inf), xhat0max=np.full(nx, inf),
                        C_umin=np.zeros(nu * Hp), C_umax=np.zeros(nu * Hp), C_dumin=np.zeros(self.nDU),
                        C_dumax=np.zeros(self.nDU), C_ymin=np.ones(ny * Hp), C_ymax=np.ones(ny * Hp),
                        c_xmin=np.ones(nx), c_xmax=np.ones(nx))
        self.Ztilde = np.zeros(self.n)
        self.lastu0 = np.zeros(nu)
        self.solved_once = False
        self._build()

    # ---- everything that depends on the model / weights (constructor and setmodel!) ----
    def _build(self):
        estim, m = self.estim, self.estim.model
        nu, ny, nd, nx, Hp, Hc, neps, nZ = m.nu, m.ny, m.nd, estim.nxhat, self.Hp, self.Hc, self.neps, self.nZ
        self.M_Hp = np.diag(np.tile(self.Mwt, Hp))
        self.L_Hp = np.diag(np.tile(self.Lwt, Hp))
        Nt = np.zeros((self.nDU + neps, self.nDU + neps))
        Nt[:self.nDU, :self.nDU] = np.diag(np.tile(self.Nwt, Hc))
        if neps:
            Nt[-1, -1] = self.Cwt
        self.Ntilde_Hc = Nt
        PDu = np.hstack([np.eye(self.nDU), np.zeros((self.nDU, self.nX))])      # init_ZtoΔU, construct.jl:743-757
        self.Pu, self.Tu = init_ZtoU(nu, Hp, Hc, self.nb, nZ)                   # zero columns for X̂0
        self.E, self.J, self.ex = init_predmat_ms(estim, Hp, Hc)
        self.ES, self.GS, self.JS, self.KS, self.VS, self.BS = init_defectmat(estim, Hp, Hc, self.nb)
        z1 = lambda r: np.zeros((r, 1))
        if neps:
            self.Ptilde_u = np.hstack([self.Pu, z1(nu * Hp)])
            self.Ptilde_Du = np.block([[PDu, z1(self.nDU)], [np.zeros((1, nZ)), np.ones((1, 1))]])
            self.Etilde = np.hstack([self.E, z1(ny * Hp)])
            self.etilde_x = np.hstack([self.ex, z1(nx)])
            self.Aeq = np.hstack([self.ES, z1(nx * Hp)])                         # augmentdefect
        else:
            self.Ptilde_u, self.Ptilde_Du, self.Etilde, self.etilde_x, self.Aeq = self.Pu, PDu, self.E, self.ex, self.ES
        self._PDu = PDu
        self.Uop, self.Yop, self.Dop = np.tile(m.uop, Hp), np.tile(m.yop, Hp), np.tile(m.dop, Hp)
        self.Htilde = 2 * (self.Etilde.T @ self.M_Hp @ self.Etilde + self.Ptilde_Du.T @ self.Ntilde_Hc @ self.Ptilde_Du
                           + self.Ptilde_u.T @ self.L_Hp @ self.Ptilde_u)
        self._rebuild_constraints()

    def _rebuild_constraints(self):
        c, neps, nZ, nDU, nX = self.con, self.neps, self.nZ, self.nDU, self.nX
        nx = self.estim.nxhat
        col = lambda v: v.reshape(-1, 1)
        rel = lambda M, cmin, cmax: ((-np.hstack([M, col(cmin)]), np.hstack([M, -col(cmax)])) if neps else (-M, M))
        A_Umin, A_Umax = rel(self.Pu, c["C_umin"], c["C_umax"])
        A_DUmin, A_DUmax = rel(self._PDu, c["C_dumin"], c["C_dumax"])
        A_Ymin, A_Ymax = rel(self.E, c["C_ymin"], c["C_ymax"])
        A_xmin, A_xmax = rel(self.ex, c["c_xmin"], c["c_xmax"])
        Zmin, Zmax = np.full(self.n, -np.inf), np.full(self.n, np.inf)
        ib = nDU + nX - nx
        if neps:
            Zmin[-1] = 0.0
            hmin, hmax = c["C_dumin"] == 0, c["C_dumax"] == 0
            Zmin[:nDU][hmin], Zmax[:nDU][hmax] = c["DUmin"][hmin], c["DUmax"][hmax]
            xmn, xmx = c["c_xmin"] == 0, c["c_xmax"] == 0                        # boxconstraint_terminal!
            Zmin[ib:ib + nx][xmn], Zmax[ib:ib + nx][xmx] = c["xhat0min"][xmn], c["xhat0max"][xmx]
        else:
            Zmin[:nDU], Zmax[:nDU] = c["DUmin"], c["DUmax"]
            Zmin[ib:ib + nx], Zmax[ib:ib + nx] = c["xhat0min"], c["xhat0max"]
        fin = np.isfinite
        i_DUmin, i_DUmax = fin(c["DUmin"]) & ~fin(Zmin[:nDU]), fin(c["DUmax"]) & ~fin(Zmax[:nDU])
        i_xmin, i_xmax = fin(c["xhat0min"]) & ~fin(Zmin[ib:ib + nx]), fin(c["xhat0max"]) & ~fin(Zmax[ib:ib + nx])
        self.i_b = np.concatenate([fin(c["U0min"]), fin(c["U0max"]), i_DUmin, i_DUmax, fin(c["Y0min"]), fin(c["Y0max"]),
                                   i_xmin, i_xmax])
        self.A = np.vstack([A_Umin, A_Umax, A_DUmin, A_DUmax, A_Ymin, A_Ymax, A_xmin, A_xmax])
        self.Zmin, self.Zmax = Zmin, Zmax

    def setconstraint(self, umin=None, umax=None, dumin=None, dumax=None, ymin=None, ymax=None, xhatmin=None,
                      xhatmax=None, c_umin=None, c_umax=None, c_dumin=None, c_dumax=None, c_ymin=None, c_ymax=None,
                      c_xhatmin=None, c_xhatmax=None):
        c, Hp, Hc, m = self.con, self.Hp, self.Hc, self.model
        nu, ny, nx = m.nu, m.ny, self.estim.nxhat
        v = lambda a, n: np.asarray(a, float).reshape(n)
        if umin is not None: c["U0min"] = np.tile(v(umin, nu), Hp) - self.Uop
        if umax is not None: c["U0max"] = np.tile(v(umax, nu), Hp) - self.Uop
        if dumin is not None: c["DUmin"] = np.tile(v(dumin, nu), Hc)
        if dumax is not None: c["DUmax"] = np.tile(v(dumax, nu), Hc)
        if ymin is not None: c["Y0min"] = np.tile(v(ymin, ny), Hp) - self.Yop
        if ymax is not None: c["Y0max"] = np.tile(v(ymax, ny), Hp) - self.Yop
        if xhatmin is not None: c["xhat0min"] = v(xhatmin, nx) - self.estim.xophat
        if xhatmax is not None: c["xhat0max"] = v(xhatmax, nx) - self.estim.xophat
        soft = dict(C_umin=(c_umin, nu, Hp), C_umax=(c_umax, nu, Hp), C_dumin=(c_dumin, nu, Hc), C_dumax=(c_dumax, nu, Hc),
                    C_ymin=(c_ymin, ny, Hp), C_ymax=(c_ymax, ny, Hp), c_xmin=(c_xhatmin, nx, 1), c_xmax=(c_xhatmax, nx, 1))
        for k, (val, n, reps) in soft.items():
            if val is not None:
                if not self.neps:
                    raise ValueError("Slack variable weight Cwt must be finite to set softness parameters")
                if self.solved_once:
                    raise RuntimeError("Cannot set softness parameters after calling moveinput!")
                c[k] = np.tile(v(val, n), reps)
        self._rebuild_constraints()
        return self

    def moveinput(self, ry=None, d=(), Dhat=None, Rhat_y=None, Rhat_u=None):
        m, Hp = self.model, self.Hp
        ry = m.yop if ry is None else np.asarray(ry, float).reshape(-1)
        d = np.asarray(d, float).reshape(-1)
        Dhat = np.tile(d, Hp) if Dhat is None else np.asarray(Dhat, float)
        self.Rhat_y = np.tile(ry, Hp) if Rhat_y is None else np.asarray(Rhat_y, float)
        self.Rhat_u = self.Uop if Rhat_u is None else np.asarray(Rhat_u, float)
        d0, Dhat0 = d - m.dop, Dhat - self.Dop
        x0 = self.estim.xhat0
        # initpred! (execute.jl:247-277): K = V = G = B = 0 for MultipleShooting
        self.F = self.J @ Dhat0 if m.nd else np.zeros(m.ny * Hp)
        Cy = self.F + self.Yop - self.Rhat_y
        self.Tu_lastu0 = self.Tu @ self.lastu0
        Cu = self.Tu_lastu0 + self.Uop - self.Rhat_u
        self.qtilde = 2 * ((self.M_Hp @ self.Etilde).T @ Cy + (self.L_Hp @ self.Ptilde_u).T @ Cu)
        self.r = Cy @ self.M_Hp @ Cy + Cu @ self.L_Hp @ Cu
        # linconstraint! (fx̂ = 0: every terminal matrix but ex̂ is zero) and linconstrainteq!
        c = self.con
        nx = self.estim.nxhat
        b = np.concatenate([-c["U0min"] + self.Tu_lastu0, c["U0max"] - self.Tu_lastu0, -c["DUmin"], c["DUmax"],
                            -c["Y0min"] + self.F, c["Y0max"] - self.F, -c["xhat0min"], c["xhat0max"]])
        FS = self.BS + self.KS @ x0 + self.VS @ self.lastu0
        if m.nd:
            FS = FS + self.GS @ d0 + self.JS @ Dhat0
        self.FS = FS
        # warm start (transcription.jl:1089-1102), the fallback on solver error
        nu, nDU, nX = m.nu, self.nDU, self.nX
        Zs = np.zeros(self.n)
        Zs[:nDU - nu] = self.Ztilde[nu:nDU]
        Zs[nDU:nDU + nX - nx] = self.Ztilde[nDU + nx:nDU + nX]
        Zs[nDU + nX - nx:nDU + nX] = self.Ztilde[nDU + nX - nx:nDU + nX]
        if self.neps:
            Zs[-1] = self.Ztilde[-1]
        sol = _qp.solve_qp_eq(self.Htilde, self.qtilde, self.A[self.i_b], b[self.i_b], self.Aeq, -FS, self.Zmin, self.Zmax)
        self.last_qp, self.last_status, self.solved_once = sol, sol["status"], True
        self.Ztilde = Zs if sol["status"] == _qp.INFEASIBLE else sol["z"].copy()
        u = self.Ztilde[:nu] + self.lastu0 + m.uop
        self.lastu0 = u - m.uop
        return u

    def getinfo(self):
        m, Z = self.model, self.Ztilde
        U0 = self.Ptilde_u @ Z + self.Tu_lastu0
        Y0 = self.Etilde @ Z + self.F
        xend = self.etilde_x @ Z
        Ybar, Ubar = Y0 + self.Yop - self.Rhat_y, U0 + self.Uop - self.Rhat_u
        DUt = self.Ptilde_Du @ Z
        J = Ybar @ self.M_Hp @ Ybar + DUt @ self.Ntilde_Hc @ DUt + Ubar @ self.L_Hp @ Ubar
        return dict(DU=Z[:self.nDU].copy(), X0=Z[self.nDU:self.nDU + self.nX].copy(), eps=(Z[-1] if self.neps else 0.0), J=J,
                    U=U0 + self.Uop, u=(U0 + self.Uop)[:m.nu], Yhat=Y0 + self.Yop, xhatend=xend + self.estim.xophat,
                    J_quad=0.5 * Z @ self.Htilde @ Z + self.qtilde @ Z + self.r)

    def setmodel(self, model=None, Mwt=None, Nwt=None, Lwt=None, **kw):
        """setmodel! for the MultipleShooting controller (src/controller/execute.jl:621-790; diagonal weights)."""
        m = self.estim.model
        uop_old, xop_old = m.uop.copy(), self.estim.xophat.copy()
        Uop_old, Yop_old = self.Uop.copy(), self.Yop.copy()
        self.estim.setmodel(m if model is None else model, **kw)
        m = self.estim.model
        if Mwt is not None: self.Mwt = np.asarray(Mwt, float).reshape(m.ny)
        if Nwt is not None: self.Nwt = np.asarray(Nwt, float).reshape(m.nu)
        if Lwt is not None: self.Lwt = np.asarray(Lwt, float).reshape(m.nu)
        c = self.con
        Uop, Yop = np.tile(m.uop, self.Hp), np.tile(m.yop, self.Hp)
        c["U0min"], c["U0max"] = c["U0min"] + Uop_old - Uop, c["U0max"] + Uop_old - Uop
        c["Y0min"], c["Y0max"] = c["Y0min"] + Yop_old - Yop, c["Y0max"] + Yop_old - Yop
        c["xhat0min"], c["xhat0max"] = c["xhat0min"] + xop_old - self.estim.xophat, c["xhat0max"] + xop_old - self.estim.xophat
        self.lastu0 = self.lastu0 + uop_old - m.uop
        self._build()
        return self

    def preparestate(self, ym, d=()):
        return self.estim.preparestate(ym, d)

    def updatestate(self, u, ym, d=()):
        return self.estim.updatestate(u, ym, d)
