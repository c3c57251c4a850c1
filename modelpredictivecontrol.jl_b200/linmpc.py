"""``LinMPC``: host-side mirror of the reference controller API for a BATCH of controllers.

Same names and argument meaning as the reference (src/controller/linmpc.jl:229-316,
``setconstraint!`` construct.jl:324-559, ``moveinput!`` execute.jl:59-80, ``getinfo`` :145-198,
``preparestate!/updatestate!`` :523-555), with a leading batch axis on every array.  The per-period
work is one call into libbmpc.so; there is no Python/CPU implementation of the step.
"""
import numpy as np

from .batch import BatchLinMPC
from .host import LinModel, ManualEstimator, SteadyKalmanFilter, _b, expand_softness, move_blocking

DEFAULT_HP0, DEFAULT_HC, DEFAULT_MWT, DEFAULT_NWT, DEFAULT_LWT, DEFAULT_CWT = 10, 2, 1.0, 0.1, 0.0, 1e5


class LinMPC:
    def __init__(self, model_or_estim, Hp=None, Hc=DEFAULT_HC, Mwt=None, Nwt=None, Lwt=None, Cwt=DEFAULT_CWT,
                 Wy=None, Wu=None, Wd=None, Wr=None, transcription="singleshooting", device=0, team=0, max_iter=0, tol=0.0,
                 fused_estimator=False, **estim_kwargs):
        estim = model_or_estim if hasattr(model_or_estim, "Ahat") else SteadyKalmanFilter(model_or_estim, **estim_kwargs)
        model = estim.model
        self.estim, self.model = estim, model
        N, nu, ny, nd = model.N, model.nu, model.ny, model.nd
        if Hp is None:
            # default_Hp (construct.jl:569-591): 10 + the number of (near-)zero poles, the estimate of the plant's
            # dead time; one horizon for the whole batch, so the estimate must agree across the instances
            nk = np.sum(np.abs(np.linalg.eigvals(model.A)) < 1e-3, axis=-1)
            if nk.min() != nk.max():
                raise ValueError("default Hp: the instances have different delay estimates; pass Hp explicitly")
            Hp = DEFAULT_HP0 + int(nk.max())
        # validate_weights (construct.jl:105-123): ArgumentError / DimensionMismatch -> ValueError
        if Hp < 1:
            raise ValueError("Prediction horizon Hp should be >= 1")
        if np.ndim(Cwt) != 0:
            raise ValueError("Cwt should be a real scalar")
        if Cwt < 0:
            raise ValueError("Cwt weight should be >= 0")
        self.nb = move_blocking(Hp, Hc)
        self.Hp, self.Hc = Hp, len(self.nb)
        w = lambda v, dflt, n: np.full(n, dflt) if v is None else np.asarray(v, dtype=np.float64).reshape(n)
        self.Mwt, self.Nwt, self.Lwt = w(Mwt, DEFAULT_MWT, ny), w(Nwt, DEFAULT_NWT, nu), w(Lwt, DEFAULT_LWT, nu)
        if (self.Mwt < 0).any() or (self.Nwt < 0).any() or (self.Lwt < 0).any():
            raise ValueError("weights should be nonnegative")
        self.Cwt = float(Cwt)
        # transcription: "singleshooting" (default) or "multipleshooting".  Both are solved as the condensed QP (same optimum
        # for a LinModel); with MultipleShooting ``Ztilde`` is assembled in the reference's layout [ΔU; X̂0; ε]
        self.transcription = str(transcription).lower().replace("_", "")
        if self.transcription not in ("singleshooting", "multipleshooting"):
            raise ValueError("transcription must be 'singleshooting' or 'multipleshooting'")
        self.batch = BatchLinMPC(N, nu, ny, estim.nxhat, Hp, self.nb, nd=nd, Cwt=Cwt, device=device, team=team,
                                 max_iter=max_iter, tol=tol)
        b = self.batch
        self._W = (Wy, Wu, Wd, Wr)
        self._push_model(first=True)
        # InternalModel: stochastic output predictions Ŷs = Ks x̂s + Ps ŷs enter F every period (predictstoch!)
        self._stoch = estim.stochpred(Hp) if hasattr(estim, "stochpred") else None
        self.Uop, self.Yop = np.tile(model.uop, (1, Hp)), np.tile(model.yop, (1, Hp))
        inf = np.inf
        self.con = dict(U0min=np.full((N, nu * Hp), -inf), U0max=np.full((N, nu * Hp), inf),
                        DUmin=np.full((N, nu * self.Hc), -inf), DUmax=np.full((N, nu * self.Hc), inf),
                        Y0min=np.full((N, ny * Hp), -inf), Y0max=np.full((N, ny * Hp), inf),
                        xhat0min=np.full((N, estim.nxhat), -inf), xhat0max=np.full((N, estim.nxhat), inf),
                        Wmin=np.full((N, self.nw * (Hp + 1)), -inf), Wmax=np.full((N, self.nw * (Hp + 1)), inf))
        self.soft = dict(C_umin=np.zeros(nu * Hp), C_umax=np.zeros(nu * Hp), C_dumin=np.zeros(nu * self.Hc),
                         C_dumax=np.zeros(nu * self.Hc), C_ymin=np.ones(ny * Hp), C_ymax=np.ones(ny * Hp),
                         c_xmin=np.ones(estim.nxhat), c_xmax=np.ones(estim.nxhat))
        self.soft_w = dict(C_wmin=np.ones(self.nw * (Hp + 1)), C_wmax=np.ones(self.nw * (Hp + 1)))
        self._solved = False
        # fused_estimator: the SteadyKalmanFilter runs INSIDE the step kernel (correct before moveinput!, predict after
        # it) and x̂0 lives in the handle: preparestate only notes ym, updatestate does nothing, one launch per period.
        self.fused_estimator = bool(fused_estimator)
        if self.fused_estimator:
            if not hasattr(estim, "Khat"):
                raise ValueError("fused_estimator needs a SteadyKalmanFilter or a KalmanFilter")
            tv = hasattr(estim, "P0hat")  # time-varying KalmanFilter: the gain comes from the covariance recursion
            b.set_estimator(estim.Ahat, estim.Buhat, estim.Cmhat, None if tv else estim.Khat, estim.Bdhat if nd else None,
                            estim.Ddmhat if nd else None, estim.fophat - estim.xophat)
            if tv:
                rep = lambda M: np.broadcast_to(M, (N,) + M.shape)
                b.set_estimator_cov(estim.Phat, rep(estim.Qhat), rep(estim.Rhat))
            b.set_state(estim.xhat0)
            self._y0m = None
        self._push()

    def _push_model(self, first=False):
        """Route A: the augmented model, the diagonal weights, the operating points and the custom-constraint matrices go to
        the handle; prediction matrices, Hessian and the custom rows' matrix are (re)built on the device."""
        estim, model, b, Hp = self.estim, self.estim.model, self.batch, self.Hp
        nu, ny, nd = model.nu, model.ny, model.nd
        b.set_model(estim.Ahat, estim.Buhat, estim.Chat, estim.Bdhat if nd else None, estim.Ddhat if nd else None,
                    estim.fophat - estim.xophat, np.tile(self.Mwt, Hp), np.tile(self.Nwt, self.Hc),
                    np.tile(self.Lwt, Hp))
        b.set_oppoints(model.uop, model.yop)
        # custom linear constraints Wy, Wu, Wd, Wr (validate_custom_lincon, construct.jl:666-695): nw rows, shared by the batch
        Wy, Wu, Wd, Wr = self._W
        given = [np.atleast_2d(np.asarray(W, float)) for W in (Wy, Wu, Wd, Wr) if W is not None]
        self.nw = given[0].shape[0] if given else 0
        if any(g.shape[0] != self.nw for g in given):
            raise ValueError("Wy, Wu, Wd, Wr must have the same number of rows")
        if self.nw and first:
            z = lambda W, nc: np.zeros((self.nw, nc)) if W is None else np.atleast_2d(np.asarray(W, float)).reshape(self.nw, nc)
            b.set_custom(self.nw, z(Wy, ny), z(Wu, nu), z(Wd, nd) if nd else None, z(Wr, ny), estim.Chat,
                         estim.Ddhat if nd else None, model.dop if nd else None)

    def setmodel(self, model=None, Mwt=None, Nwt=None, Lwt=None):
        """Batched ``setmodel!`` (reference src/controller/execute.jl:621-790): new plant models and / or diagonal weights
        at run time.  Z̃ is kept; u0(k-1) and the deviation-form bounds are re-expressed around the new operating points
        (:757-776); prediction matrices and Hessian are rebuilt ON THE DEVICE (bmpc_set_model)."""
        if self.nw:
            raise NotImplementedError("setmodel with custom linear constraints (the reference drops them too: Appendix C-2)")
        if self.fused_estimator:
            raise NotImplementedError("setmodel with the fused estimator: re-create the controller")
        m_old = self.estim.model
        N, Hp = m_old.N, self.Hp
        uop_old, yop_old, xop_old = m_old.uop.copy(), m_old.yop.copy(), self.estim.xophat.copy()
        if model is not None:
            self.estim.setmodel(model)
            self.model = model
        m = self.estim.model
        chk = lambda w, n, name: np.asarray(w, dtype=np.float64).reshape(n)
        if Mwt is not None: self.Mwt = chk(Mwt, m.ny, "Mwt")
        if Nwt is not None: self.Nwt = chk(Nwt, m.nu, "Nwt")
        if Lwt is not None: self.Lwt = chk(Lwt, m.nu, "Lwt")
        if (self.Mwt < 0).any() or (self.Nwt < 0).any() or (self.Lwt < 0).any():
            raise ValueError("weights should be nonnegative")
        c = self.con
        Uop_new, Yop_new = np.tile(m.uop, (1, Hp)), np.tile(m.yop, (1, Hp))
        for k, shift in (("U0min", self.Uop - Uop_new), ("U0max", self.Uop - Uop_new), ("Y0min", self.Yop - Yop_new),
                         ("Y0max", self.Yop - Yop_new), ("xhat0min", xop_old - self.estim.xophat),
                         ("xhat0max", xop_old - self.estim.xophat)):
            c[k] = c[k] + shift
        self.batch.lastu0[:] = self.batch.lastu0 + uop_old - m.uop
        self.Uop, self.Yop = Uop_new, Yop_new
        self._push_model()
        self._push()
        return self

    @property
    def Ztilde(self):
        if self.transcription == "multipleshooting" and self._solved:
            b = self.batch
            Z = b.Ztilde
            return np.concatenate([Z[:, :b.nDU], b.get_states(), Z[:, b.nDU:]], axis=1)
        return self.batch.Ztilde

    @property
    def lastu0(self):
        return self.batch.lastu0

    def _push(self):
        c = self.con
        if self.nw:
            sw = self.soft_w if self.batch.neps else dict(C_wmin=None, C_wmax=None)
            self.batch.set_custom_bounds(c["Wmin"], c["Wmax"], sw["C_wmin"], sw["C_wmax"])
        self.batch.set_constraints(c["U0min"], c["U0max"], c["DUmin"], c["DUmax"], c["Y0min"], c["Y0max"],
                                   c["xhat0min"], c["xhat0max"], self.soft if self.batch.neps else None)

    def setconstraint(self, umin=None, umax=None, dumin=None, dumax=None, ymin=None, ymax=None, xhatmin=None,
                      xhatmax=None, Umin=None, Umax=None, DUmin=None, DUmax=None, Ymin=None, Ymax=None,
                      c_umin=None, c_umax=None, c_dumin=None, c_dumax=None, c_ymin=None, c_ymax=None,
                      c_xhatmin=None, c_xhatmax=None, wmin=None, wmax=None, Wmin=None, Wmax=None, c_wmin=None,
                      c_wmax=None, C_umin=None, C_umax=None, C_dumin=None, C_dumax=None, C_ymin=None, C_ymax=None,
                      C_wmin=None, C_wmax=None):
        """``setconstraint!`` (src/controller/construct.jl:324-559): bounds are per instance ((N, len) or (len,)), softness
        is shared.  Lower-case keywords hold for every sample of the horizon, capitalised ones give the whole horizon."""
        N, Hp, Hc = self.model.N, self.Hp, self.Hc
        nu, ny, nx = self.model.nu, self.model.ny, self.estim.nxhat
        c = self.con
        rep = lambda v, n, k: np.tile(_b(v, N, (n,), strict=True), (1, k))
        if Umin is None and umin is not None: c["U0min"] = rep(umin, nu, Hp) - self.Uop
        elif Umin is not None: c["U0min"] = _b(Umin, N, (nu * Hp,)) - self.Uop
        if Umax is None and umax is not None: c["U0max"] = rep(umax, nu, Hp) - self.Uop
        elif Umax is not None: c["U0max"] = _b(Umax, N, (nu * Hp,)) - self.Uop
        if DUmin is None and dumin is not None: c["DUmin"] = rep(dumin, nu, Hc)
        elif DUmin is not None: c["DUmin"] = _b(DUmin, N, (nu * Hc,)).copy()
        if DUmax is None and dumax is not None: c["DUmax"] = rep(dumax, nu, Hc)
        elif DUmax is not None: c["DUmax"] = _b(DUmax, N, (nu * Hc,)).copy()
        if Ymin is None and ymin is not None: c["Y0min"] = rep(ymin, ny, Hp) - self.Yop
        elif Ymin is not None: c["Y0min"] = _b(Ymin, N, (ny * Hp,)) - self.Yop
        if Ymax is None and ymax is not None: c["Y0max"] = rep(ymax, ny, Hp) - self.Yop
        elif Ymax is not None: c["Y0max"] = _b(Ymax, N, (ny * Hp,)) - self.Yop
        nw = self.nw
        if (wmin is not None or wmax is not None or Wmin is not None or Wmax is not None) and not nw:
            raise ValueError("custom bounds need the Wy / Wu / Wd / Wr matrices of the constructor")
        if Wmin is None and wmin is not None: c["Wmin"] = rep(wmin, nw, Hp + 1)            # construct.jl:410-418
        elif Wmin is not None: c["Wmin"] = _b(Wmin, N, (nw * (Hp + 1),)).copy()
        if Wmax is None and wmax is not None: c["Wmax"] = rep(wmax, nw, Hp + 1)
        elif Wmax is not None: c["Wmax"] = _b(Wmax, N, (nw * (Hp + 1),)).copy()
        if xhatmin is not None: c["xhat0min"] = _b(xhatmin, N, (nx,)) - self.estim.xophat
        if xhatmax is not None: c["xhat0max"] = _b(xhatmax, N, (nx,)) - self.estim.xophat
        # softness: sizes and signs are checked before anything is stored
        ecr = dict(C_umin=expand_softness(c_umin, C_umin, nu, Hp, "umin"), C_umax=expand_softness(c_umax, C_umax, nu, Hp, "umax"),
                   C_dumin=expand_softness(c_dumin, C_dumin, nu, Hc, "dumin"), C_dumax=expand_softness(c_dumax, C_dumax, nu, Hc, "dumax"),
                   C_ymin=expand_softness(c_ymin, C_ymin, ny, Hp, "ymin"), C_ymax=expand_softness(c_ymax, C_ymax, ny, Hp, "ymax"),
                   c_xmin=expand_softness(c_xhatmin, None, nx, 1, "xhatmin"), c_xmax=expand_softness(c_xhatmax, None, nx, 1, "xhatmax"))
        ecr_w = dict(C_wmin=expand_softness(c_wmin, C_wmin, nw, Hp + 1, "wmin"), C_wmax=expand_softness(c_wmax, C_wmax, nw, Hp + 1, "wmax"))
        if any(v is not None for v in list(ecr.values()) + list(ecr_w.values())):
            if not self.batch.neps:
                raise ValueError("Slack variable weight Cwt must be finite to set softness parameters")
            if self._solved:
                raise RuntimeError("Cannot set softness parameters after calling moveinput!")
            for k, v in ecr.items():
                if v is not None:
                    self.soft[k] = v
            for k, v in ecr_w.items():
                if v is not None:
                    self.soft_w[k] = v
        self._push()
        return self

    def initstate(self, u, ym, d=None):
        """``initstate!(mpc, u, ym, d)`` (src/controller/execute.jl:1-13): the estimator's steady state for (u, ym, d), the
        warm start Z̃ cleared, u - uop stored as u0(k-1)."""
        if self.fused_estimator:
            raise NotImplementedError("initstate with the fused estimator: set the state with batch.set_state")
        m = self.model
        self.batch.Ztilde[:] = 0.0
        self.batch.lastu0[:] = _b(u, m.N, (m.nu,)) - m.uop
        return self.estim.initstate(u, ym, d)

    # ---- estimator pass-throughs ----
    def preparestate(self, ym, d=None):
        if self.fused_estimator:
            m = self.model
            self._y0m = _b(ym, m.N, (len(self.estim.i_ym),)) - m.yop[:, self.estim.i_ym]
            return None
        return self.estim.preparestate(ym, d)

    def updatestate(self, u, ym, d=None):
        if self.fused_estimator:
            return None  # done by the step kernel
        return self.estim.updatestate(u, ym, d)

    def setstate(self, xhat):
        self.estim.setstate(xhat)
        if self.fused_estimator:
            self.batch.set_state(self.estim.xhat0)
        return self

    # ---- the hot path ----
    def moveinput(self, ry=None, d=None, lastu=None, Dhat=None, Rhat_y=None, Rhat_u=None):
        m = self.model
        N = m.N
        ry = m.yop if ry is None else _b(ry, N, (m.ny,))
        if lastu is not None:
            self.batch.lastu0[:] = _b(lastu, N, (m.nu,)) - m.uop
        d0 = Dh0 = None
        if m.nd:
            d0 = _b(d, N, (m.nd,)) - m.dop
            if Dhat is not None:
                Dh0 = _b(Dhat, N, (m.nd * self.Hp,)) - np.tile(m.dop, (1, self.Hp))
        self._solved = True
        if self.fused_estimator:
            if self._y0m is None:
                raise RuntimeError("call preparestate(ym) before moveinput with fused_estimator")
            return self.batch.step(None, ry=ry, Rhat_y=Rhat_y, Rhat_u=Rhat_u, d0=d0, Dhat0=Dh0, y0m=self._y0m).copy()
        Ys = None
        if self._stoch is not None:
            Ks, Ps = self._stoch
            Ys = self.estim.xs @ Ks.T + self.estim.ys @ Ps.T
        return self.batch.step(self.estim.xhat0, ry=ry, Rhat_y=Rhat_y, Rhat_u=Rhat_u, d0=d0, Dhat0=Dh0, Yhat_s=Ys).copy()

    def getinfo(self):
        i = self.batch.getinfo()
        out = dict(DU=i["DU"], eps=i["eps"], J=i["J"], U=i["U0"] + self.Uop, Yhat=i["Yhat0"] + self.Yop,
                   xhatend=i["xhat0end"] + self.estim.xophat, status=i["status"], iters=i["iters"])
        out["u"] = out["U"][:, :self.model.nu]
        if self.transcription == "multipleshooting":
            out["X0"] = self.batch.get_states()
        return out


def sim(mpc, steps, ry, plant=None, y_noise=None):
    """Batched ``sim!`` closed loop (reference src/plot_sim.jl:253-319): per period
    y = plant(); preparestate!; u = moveinput!(ry); updatestate! on plant and estimator."""
    plant = plant or LinModel(mpc.model.A, mpc.model.Bu, mpc.model.C, N=mpc.model.N, uop=mpc.model.uop,
                              yop=mpc.model.yop, xop=mpc.model.xop, fop=mpc.model.fop)
    N = mpc.model.N
    Y, U = np.zeros((steps, N, mpc.model.ny)), np.zeros((steps, N, mpc.model.nu))
    for k in range(steps):
        r = ry(k) if callable(ry) else ry
        y = plant.evaloutput()
        if y_noise is not None:
            y = y + y_noise[k]
        mpc.preparestate(y)
        u = mpc.moveinput(r)
        Y[k], U[k] = y, u
        plant.updatestate(u)
        mpc.updatestate(u, y)
    return dict(Y=Y, U=U)
