"""numpy fp64 restatement of the reference's ``LinMPC`` hot path (TEST INFRASTRUCTURE ONLY).

Follows, function by function, JuliaControl/ModelPredictiveControl.jl v2.11.0 @ 56aa83a
(citations relative to /root/reference).  Unicode names are ASCII-ised:
``x̂ -> xhat``, ``Ẽ -> Etilde``, ``ΔU -> DU``, ``ϵ -> eps``, ``Z̃ -> Ztilde``.
Vectors over the horizon are stacked by time step, matrices are column-major in the
reference; numpy arrays here are ordinary 2-D arrays with the same (row, col) meaning.
"""
from __future__ import annotations

import dataclasses
import numpy as np
import scipy.linalg

from . import qp as _qp

# src/general.jl:1-7
DEFAULT_HP0, DEFAULT_HC, DEFAULT_MWT, DEFAULT_NWT, DEFAULT_LWT, DEFAULT_CWT = 10, 2, 1.0, 0.1, 0.0, 1e5


# --------------------------------------------------------------------------------------
# LinModel  (src/model/linmodel.jl:1-66, direct-matrix constructor :252-253, setop! src/sim_model.jl:101-126)
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class LinModel:
    A: np.ndarray
    Bu: np.ndarray
    C: np.ndarray
    Bd: np.ndarray | None = None
    Dd: np.ndarray | None = None
    Ts: float = 1.0
    uop: np.ndarray | None = None
    yop: np.ndarray | None = None
    dop: np.ndarray | None = None
    xop: np.ndarray | None = None
    fop: np.ndarray | None = None

    def __post_init__(self):
        self.A = np.atleast_2d(np.asarray(self.A, float))
        self.nx = self.A.shape[0]
        self.Bu = np.asarray(self.Bu, float).reshape(self.nx, -1)
        self.nu = self.Bu.shape[1]
        self.C = np.asarray(self.C, float).reshape(-1, self.nx)
        self.ny = self.C.shape[0]
        self.Bd = np.zeros((self.nx, 0)) if self.Bd is None else np.asarray(self.Bd, float).reshape(self.nx, -1)
        self.nd = self.Bd.shape[1]
        self.Dd = np.zeros((self.ny, self.nd)) if self.Dd is None else np.asarray(self.Dd, float).reshape(self.ny, self.nd)
        z = lambda v, n: np.zeros(n) if v is None else np.asarray(v, float).reshape(n)
        self.uop, self.yop, self.dop = z(self.uop, self.nu), z(self.yop, self.ny), z(self.dop, self.nd)
        self.xop, self.fop = z(self.xop, self.nx), z(self.fop, self.nx)
        self.x0 = np.zeros(self.nx)

    # plant simulation (src/sim_model.jl:239-277): x0(k+1) = A x0 + Bu u0 + Bd d0 + fop - xop
    def evaloutput(self, d=()):
        d0 = np.asarray(d, float).reshape(self.nd) - self.dop
        return self.C @ self.x0 + self.Dd @ d0 + self.yop

    def updatestate(self, u, d=()):
        u0 = np.asarray(u, float) - self.uop
        d0 = np.asarray(d, float).reshape(self.nd) - self.dop
        self.x0 = self.A @ self.x0 + self.Bu @ u0 + self.Bd @ d0 + self.fop - self.xop
        return self.x0 + self.xop

    def setstate(self, x):
        self.x0 = np.asarray(x, float) - self.xop


def zoh_first_order(gain, tau, Ts, b=1.0):
    """ZOH discretisation of gain/(tau s+1) realised as  xdot = -x/tau + b u,  y = (gain/(tau b)) x.

    The reference builds such plants with ControlSystemsBase ``tf -> ss -> c2d`` (src/model/
    linmodel.jl:148-239, third-party realisation, out of scope); only the (b, c) split of the
    realisation is a free choice and it matters solely through the Kalman gain.
    """
    a = np.exp(-Ts / tau)
    return np.array([[a]]), np.array([[b * tau * (1 - a)]]), np.array([[gain / (tau * b)]])


# --------------------------------------------------------------------------------------
# Estimator construction  (src/estimator/construct.jl)
# --------------------------------------------------------------------------------------
def init_integrators(nint, ny):
    """src/estimator/construct.jl:226-251."""
    nint = np.zeros(ny, int) if np.isscalar(nint) and nint == 0 else np.asarray(nint, int).reshape(-1)
    if nint.size != ny:
        raise ValueError("nint length != n outputs")
    if (nint < 0).any():
        raise ValueError("nint values should be >= 0")
    nx = int(nint.sum())
    A, C = np.zeros((nx, nx)), np.zeros((ny, nx))
    i_A = 0
    for i in range(ny):
        k = int(nint[i])
        if k:
            A[i_A:i_A + k, i_A:i_A + k] = np.eye(k) + np.eye(k, k=-1)
            C[i, i_A + k - 1] = 1.0
            i_A += k
    return A, C, nint


def init_estimstoch(model, i_ym, nint_u, nint_ym):
    """src/estimator/construct.jl:172-185 (+ stoch_ym2y :197-209)."""
    nu, ny, nym = model.nu, model.ny, len(i_ym)
    As_u, Cs_u, nint_u = init_integrators(nint_u, nu)
    As_ym, Cs_ym, nint_ym = init_integrators(nint_ym, nym)
    Cs_y = np.zeros((ny, Cs_ym.shape[1]))
    Cs_y[list(i_ym), :] = Cs_ym
    nxs_u, nxs_y = As_u.shape[0], As_ym.shape[0]
    As = np.block([[As_u, np.zeros((nxs_u, nxs_y))], [np.zeros((nxs_y, nxs_u)), As_ym]])
    Cs_u = np.hstack([Cs_u, np.zeros((nu, nxs_y))])
    Cs_y = np.hstack([np.zeros((ny, nxs_u)), Cs_y])
    return As, Cs_u, Cs_y, nint_u, nint_ym


def observable(A, C):
    n = A.shape[0]
    O = np.vstack([C @ np.linalg.matrix_power(A, k) for k in range(n)])
    return np.linalg.matrix_rank(O) == n


def augment_model(model, As, Cs_u, Cs_y, verify_obsv=True):
    """src/estimator/construct.jl:305-323."""
    nx, nu, nd = model.nx, model.nu, model.nd
    nxs = As.shape[0]
    Ahat = np.block([[model.A, model.Bu @ Cs_u], [np.zeros((nxs, nx)), As]])
    Buhat = np.vstack([model.Bu, np.zeros((nxs, nu))])
    Chat = np.hstack([model.C, Cs_y])
    Bdhat = np.vstack([model.Bd, np.zeros((nxs, nd))])
    Ddhat = model.Dd
    if verify_obsv and not observable(Ahat, Chat):
        raise ValueError("The augmented model is unobservable.")
    xophat = np.concatenate([model.xop, np.zeros(nxs)])
    fophat = np.concatenate([model.fop, np.zeros(nxs)])
    return Ahat, Buhat, Chat, Bdhat, Ddhat, xophat, fophat


def default_nint(model, i_ym=None, nint_u=0):
    """src/estimator/construct.jl:365-376."""
    i_ym = list(range(model.ny)) if i_ym is None else list(i_ym)
    nint_ym = [0] * len(i_ym)
    for i in range(len(i_ym)):
        nint_ym[i] = 1
        As, Cs_u, Cs_y, _, _ = init_estimstoch(model, i_ym, nint_u, nint_ym)
        Ahat, _, Chat, *_ = augment_model(model, As, Cs_u, Cs_y, verify_obsv=False)
        if not observable(Ahat, Chat):
            nint_ym[i] = 0
    return nint_ym


class StateEstimator:
    """Common data of the estimators used on the LinMPC path (src/state_estim.jl, estimator/execute.jl)."""

    direct = True

    def _init_common(self, model, i_ym, nint_u, nint_ym):
        self.model = model
        self.i_ym = list(range(model.ny)) if i_ym is None else list(i_ym)
        if nint_ym is None:
            nint_ym = default_nint(model, self.i_ym, nint_u)
        As, Cs_u, Cs_y, self.nint_u, self.nint_ym = init_estimstoch(model, self.i_ym, nint_u, nint_ym)
        self._As, self._Cs_u, self._Cs_y = As, Cs_u, Cs_y
        (self.Ahat, self.Buhat, self.Chat, self.Bdhat, self.Ddhat,
         self.xophat, self.fophat) = augment_model(model, As, Cs_u, Cs_y)
        self.nxhat = self.Ahat.shape[0]
        self.nxs = As.shape[0]
        self.Cmhat, self.Ddmhat = self.Chat[self.i_ym], self.Ddhat[self.i_ym]
        self.xhat0 = np.zeros(self.nxhat)

    def evaloutput(self, d=()):
        """src/estimator/execute.jl:287-298: yhat = Chat xhat0 + Ddhat d0 + yop."""
        d0 = np.asarray(d, float).reshape(self.model.nd) - self.model.dop
        return self.Chat @ self.xhat0 + self.Ddhat @ d0 + self.model.yop

    def setstate(self, xhat):
        self.xhat0 = np.asarray(xhat, float) - self.xophat

    def preparestate(self, ym, d=()):
        """src/estimator/execute.jl:334-345."""
        if self.direct:
            y0m = np.asarray(ym, float) - self.model.yop[self.i_ym]
            d0 = np.asarray(d, float).reshape(self.model.nd) - self.model.dop
            self.correct_estimate(y0m, d0)
        return self.xhat0 + self.xophat

    def updatestate(self, u, ym, d=()):
        """src/estimator/execute.jl:374-386."""
        y0m = np.asarray(ym, float) - self.model.yop[self.i_ym]
        d0 = np.asarray(d, float).reshape(self.model.nd) - self.model.dop
        u0 = np.asarray(u, float) - self.model.uop
        self.update_estimate(u0, y0m, d0)
        return self.xhat0 + self.xophat

    def initstate(self, u, ym, d=()):
        """init_estimate! steady-state solve, src/estimator/execute.jl:246-259."""
        y0m = np.asarray(ym, float) - self.model.yop[self.i_ym]
        d0 = np.asarray(d, float).reshape(self.model.nd) - self.model.dop
        u0 = np.asarray(u, float) - self.model.uop
        rhs_x = self.fophat - self.xophat + self.Buhat @ u0 + self.Bdhat @ d0
        rhs_y = y0m - self.Ddmhat @ d0
        M = np.vstack([np.eye(self.nxhat) - self.Ahat, self.Cmhat])
        self.xhat0 = np.linalg.lstsq(M, np.concatenate([rhs_x, rhs_y]), rcond=None)[0]
        return self.xhat0 + self.xophat

    def correct_estimate(self, y0m, d0):
        pass

    def update_estimate(self, u0, y0m, d0):
        pass

    def setmodel(self, model, Qhat=None, Rhat=None):
        """setmodel!(estim, model; Q̂, R̂) (src/estimator/execute.jl:483-544): the new LinModel's matrices and operating
        points replace the old ones (setmodel_linmodel!, :500-516), the augmented matrices are rebuilt with the same
        stochastic model, and x̂0 is re-expressed around the new x̂op (setmodel_estimator!, :524-544)."""
        old = self.model
        if (model.nu, model.ny, model.nd, model.nx) != (old.nu, old.ny, old.nd, old.nx):
            raise ValueError("model dimensions must be the same")
        if model is not old:
            for k in ("A", "Bu", "C", "Bd", "Dd", "uop", "yop", "dop", "xop", "fop"):
                setattr(old, k, np.array(getattr(model, k), dtype=float, copy=True))
        self._setmodel_estimator(Qhat, Rhat)
        return self

    def _setmodel_estimator(self, Qhat, Rhat):
        Ah, Bu, Ch, Bd, Dd, xop, fop = augment_model(self.model, self._As, self._Cs_u, self._Cs_y, verify_obsv=False)
        self.Ahat, self.Buhat, self.Chat, self.Bdhat, self.Ddhat = Ah, Bu, Ch, Bd, Dd
        self.Cmhat, self.Ddmhat = Ch[self.i_ym], Dd[self.i_ym]
        xhat = self.xhat0 + self.xophat           # x̂ with the old operating point
        self.xophat, self.fophat = xop, fop
        self.xhat0 = xhat - self.xophat
        if Qhat is not None:
            self.Qhat = np.atleast_2d(np.asarray(Qhat, float))
        if Rhat is not None:
            self.Rhat = np.atleast_2d(np.asarray(Rhat, float))


class SteadyKalmanFilter(StateEstimator):
    """src/estimator/kalman.jl:163-227 (gain) and :284-309 (correct / predict).

    ``ControlSystemsBase.kalman(Discrete, Ahat, Chat, Qhat, Rhat; direct)`` is third-party; its
    documented result is the DARE solution P with Khat = P C'(C P C' + R)^-1 (direct=true,
    filter form) or A P C'(C P C' + R)^-1 (direct=false); pinned jointly with the rest of the
    chain by the golden doctest u = 17.577311 (ext/LinearMPCext.jl:252-261).
    """

    def __init__(self, model, i_ym=None, sigmaQ=None, sigmaR=None, nint_u=0, nint_ym=None,
                 sigmaQint_u=None, sigmaQint_ym=None, direct=True):
        self._init_common(model, i_ym, nint_u, nint_ym)
        self.direct = direct
        nym = len(self.i_ym)
        sigmaQ = np.full(model.nx, 1.0 / model.nx) if sigmaQ is None else np.asarray(sigmaQ, float)
        sigmaR = np.ones(nym) if sigmaR is None else np.asarray(sigmaR, float)
        sQu = np.ones(int(np.sum(self.nint_u))) if sigmaQint_u is None else np.asarray(sigmaQint_u, float)
        sQy = np.ones(int(np.sum(self.nint_ym))) if sigmaQint_ym is None else np.asarray(sigmaQint_ym, float)
        self.Qhat = np.diag(np.concatenate([sigmaQ, sQu, sQy]) ** 2)
        self.Rhat = np.diag(sigmaR ** 2)
        A, Cm = self.Ahat, self.Cmhat
        P = scipy.linalg.solve_discrete_are(A.T, Cm.T, self.Qhat, self.Rhat)
        S = Cm @ P @ Cm.T + self.Rhat
        K = np.linalg.solve(S.T, (P @ Cm.T).T).T
        self.Khat = K if direct else A @ K
        self.Phat = P

    def _setmodel_estimator(self, Qhat, Rhat):
        # src/estimator/kalman.jl:229-234
        raise RuntimeError("SteadyKalmanFilter does not support setmodel! (use KalmanFilter instead)")

    def correct_estimate(self, y0m, d0):
        if np.isnan(y0m).any():  # kalman.jl:248-251
            return
        vhat = y0m - (self.Cmhat @ self.xhat0 + self.Ddmhat @ d0)
        self.xhat0 = self.xhat0 + self.Khat @ vhat

    def update_estimate(self, u0, y0m, d0):
        if not self.direct:
            self.correct_estimate(y0m, d0)
        self.xhat0 = (self.Ahat @ self.xhat0 + self.Buhat @ u0 + self.Bdhat @ d0
                      + self.fophat - self.xophat)


class InternalModel(StateEstimator):
    """``InternalModel`` estimator (src/estimator/internal_model.jl:1-381) for a LinModel: the plant model gives the
    deterministic state x̂d (no augmentation: Â = A ..., matrices_internalmodel :158-162), a stochastic model
    ``stoch_ym = (As, Bs, Cs, Ds)`` of the measured outputs (default: one integrator per measured output,
    ss(I, I, I, I), :96) gives x̂s through  x̂s(k+1) = Âs x̂s + B̂s ŷs,  Âs = As - Bs Ds^-1 Cs,  B̂s = Bs Ds^-1
    (init_internalmodel :222-226);  ŷs^m = ym - ŷd^m,  ŷs^u = 0 (correct_estimate! :262-277)."""

    def __init__(self, model, i_ym=None, stoch_ym=None):
        if np.any(np.abs(np.linalg.eigvals(model.A)) >= 1):
            raise ValueError("InternalModel does not support integrating or unstable model")
        self.model = model
        self.i_ym = list(range(model.ny)) if i_ym is None else list(i_ym)
        nym, ny = len(self.i_ym), model.ny
        if stoch_ym is None:
            Asm = Bsm = Csm = Dsm = np.eye(nym)
        else:
            Asm, Bsm, Csm, Dsm = [np.atleast_2d(np.asarray(M, float)) for M in stoch_ym]
        if Csm.shape[0] != nym or Dsm.shape[0] != nym:
            raise ValueError("Stochastic model output quantity is different from measured output quantity")
        if not np.any(Dsm):
            raise ValueError("Stochastic model requires a nonzero direct transmission matrix D")
        # stoch_ym2y (src/estimator/construct.jl): rows of the unmeasured outputs are zero
        nxs = Asm.shape[0]
        self.As, self.Bs = Asm, np.zeros((nxs, ny))
        self.Cs, self.Ds = np.zeros((ny, nxs)), np.eye(ny)
        self.Bs[:, self.i_ym] = Bsm
        self.Cs[self.i_ym] = Csm
        self.Ds[np.ix_(self.i_ym, self.i_ym)] = Dsm
        self.nxs, self.nxhat = nxs, model.nx
        self.nint_u, self.nint_ym = np.zeros(model.nu, int), np.zeros(nym, int)
        self._set_matrices()
        self.Bs_hat = self.Bs @ np.linalg.inv(self.Ds)                      # init_internalmodel
        self.As_hat = self.As - self.Bs_hat @ self.Cs
        self.xhat0 = np.zeros(model.nx)
        self.xs = np.zeros(nxs)
        self.ys = np.zeros(ny)

    def _set_matrices(self):
        m = self.model
        self.Ahat, self.Buhat, self.Chat, self.Bdhat, self.Ddhat = m.A, m.Bu, m.C, m.Bd, m.Dd
        self.xophat, self.fophat = m.xop.copy(), m.fop.copy()
        self.Cmhat, self.Ddmhat = self.Chat[self.i_ym], self.Ddhat[self.i_ym]

    def _setmodel_estimator(self, Qhat, Rhat):
        xhat = self.xhat0 + self.xophat
        self._set_matrices()
        self.xhat0 = xhat - self.xophat

    def evaloutput(self, d=()):
        """internal_model.jl:357-368: ŷ = Ĉ x̂d + D̂d d0 + yop + ŷs."""
        return StateEstimator.evaloutput(self, d) + self.ys

    def correct_estimate(self, y0m, d0):
        yd = self.Chat @ self.xhat0 + self.Ddhat @ d0
        ys = np.zeros(self.model.ny)
        for k, i in enumerate(self.i_ym):
            ys[i] = y0m[k] - yd[i] if np.isfinite(y0m[k]) else 0.0
        self.ys = ys

    def update_estimate(self, u0, y0m, d0):
        self.xhat0 = self.Ahat @ self.xhat0 + self.Buhat @ u0 + self.Bdhat @ d0 + self.fophat - self.xophat
        self.xs = self.As_hat @ self.xs + self.Bs_hat @ self.ys

    def initstate(self, u, ym, d=()):
        """init_estimate! (internal_model.jl:325-343): both parts at steady state."""
        m = self.model
        u0, d0 = np.asarray(u, float) - m.uop, np.asarray(d, float).reshape(m.nd) - m.dop
        y0m = np.asarray(ym, float) - m.yop[self.i_ym]
        self.xhat0 = np.linalg.solve(np.eye(m.nx) - m.A, m.Bu @ u0 + m.Bd @ d0 + self.fophat - self.xophat)
        self.correct_estimate(y0m, d0)
        self.xs = np.linalg.solve(np.eye(self.nxs) - self.As_hat, self.Bs_hat @ self.ys)
        return self.xhat0 + self.xophat


def init_stochpred(estim, Hp):
    """src/controller/construct.jl:1254-1272: Ŷs = Ks x̂s + Ps ŷs for an InternalModel, empty otherwise."""
    ny = estim.model.ny
    if not isinstance(estim, InternalModel):
        return np.zeros((0, estim.nxs)), np.zeros((0, ny))
    As, Bs, Cs = estim.As, estim.Bs_hat, estim.Cs
    Ks, Ps = np.zeros((ny * Hp, estim.nxs)), np.zeros((ny * Hp, ny))
    Ap = np.eye(estim.nxs)
    for i in range(1, Hp + 1):
        Ms = Cs @ Ap @ Bs                      # Cs As^(i-1) B̂s
        Ap = Ap @ As
        Ks[ny * (i - 1):ny * i] = Cs @ Ap - Ms @ Cs
        Ps[ny * (i - 1):ny * i] = Ms
    return Ks, Ps


class ManualEstimator(StateEstimator):
    """src/estimator/manual.jl:60-64,150-154: the host supplies xhat through setstate!; prepare /
    update are no-ops.  This is exactly the contract of the batched C-ABI step."""

    def __init__(self, model, i_ym=None, nint_u=0, nint_ym=None):
        self._init_common(model, i_ym, nint_u, nint_ym)


# --------------------------------------------------------------------------------------
# Controller construction  (src/controller/construct.jl, src/controller/transcription.jl)
# --------------------------------------------------------------------------------------
def move_blocking(Hp, Hc):
    """src/controller/construct.jl:629-660."""
    if np.isscalar(Hc):
        nb = [1] * int(Hc)
        if Hc > 0:
            nb[-1] = Hp - Hc + 1
        return nb
    nb = [int(v) for v in Hc]
    if not all(v > 0 for v in nb):
        raise ValueError("Move blocking vector must be strictly positive integers.")
    if sum(nb) < Hp:
        nb = nb + [Hp - sum(nb)]
    elif sum(nb) > Hp:
        cs = np.cumsum(nb)
        last = int(np.flatnonzero(cs >= Hp)[0])
        nb = nb[:last + 1]
        if sum(nb) > Hp:
            nb[-1] = Hp - sum(nb[:-1])
    return nb


def validate_weights(nu, ny, Hp, Hc, M_Hp, N_Hc, L_Hp, Cwt):
    """src/controller/construct.jl:105-123 (ArgumentError / DimensionMismatch -> ValueError)."""
    if Hp < 1: raise ValueError("Prediction horizon Hp should be >= 1")
    if Hc < 1: raise ValueError("Control horizon Hc should be >= 1")
    if Hc > Hp: raise ValueError("Control horizon Hc should be <= prediction horizon Hp")
    for M, n, name, wname in ((M_Hp, ny * Hp, "M_Hp", "Mwt"), (N_Hc, nu * Hc, "N_Hc", "Nwt"), (L_Hp, nu * Hp, "L_Hp", "Lwt")):
        if M.shape != (n, n):
            raise ValueError(f"{name} size {M.shape} != ({n}, {n})")
        if np.count_nonzero(M - np.diag(np.diag(M))) == 0 and (np.diag(M) < 0).any():
            raise ValueError(f"{wname} values should be nonnegative")
        if not np.array_equal(M, M.T):
            raise ValueError(f"{name} should be hermitian")
    if np.ndim(Cwt) != 0: raise ValueError("Cwt should be a real scalar")
    if Cwt < 0: raise ValueError("Cwt weight should be >= 0")


def init_ZtoDU(nu, Hc, nZ):
    """src/controller/construct.jl:733-741."""
    P = np.zeros((nu * Hc, nZ))
    P[:, :nu * Hc] = np.eye(nu * Hc)
    return P


def init_ZtoU(nu, Hp, Hc, nb, nZ):
    """src/controller/construct.jl:792-809."""
    Pu = np.zeros((nu * Hp, nZ))
    row = 0
    for i in range(Hc):
        for _ in range(nb[i]):
            for j in range(i + 1):
                Pu[row:row + nu, nu * j:nu * (j + 1)] = np.eye(nu)
            row += nu
    Tu = np.tile(np.eye(nu), (Hp, 1))
    return Pu, Tu


def init_predmat(estim, Hp, Hc, nb):
    """LinModel + SingleShooting, src/controller/transcription.jl:115-194."""
    A, Bu, C, Bd, Dd = estim.Ahat, estim.Buhat, estim.Chat, estim.Bdhat, estim.Ddhat
    nu, nx, ny, nd = estim.model.nu, estim.nxhat, estim.model.ny, estim.model.nd
    Apow = np.zeros((Hp + 1, nx, nx))
    Apow[0] = np.eye(nx)
    for j in range(1, Hp + 1):
        Apow[j] = Apow[j - 1] @ A
    Acsum = np.cumsum(Apow, axis=0)
    S = lambda m: Acsum[m]
    jl = np.concatenate([[0], np.cumsum(nb)]).astype(int)
    kx = Apow[Hp]
    K = np.vstack([C @ Apow[j] for j in range(1, Hp + 1)])
    vx = S(Hp - 1) @ Bu
    V = np.vstack([C @ S(l) @ Bu for l in range(Hp)])
    nZ = nu * Hc
    ex = np.zeros((nx, nZ))
    E = np.zeros((Hp * ny, nZ))
    for j in range(Hc):
        cols = slice(nu * j, nu * (j + 1))
        for i in range(j, Hc):
            i_Q, m_Q, b_Q = jl[i], jl[i + 1], jl[j]
            for l in range(m_Q - i_Q):
                r0 = ny * (i_Q + l)
                E[r0:r0 + ny, cols] = C @ S(i_Q - b_Q + l) @ Bu
        ex[:, cols] = S(Hp - jl[j] - 1) @ Bu
    gx = Apow[Hp - 1] @ Bd
    G = np.zeros((Hp * ny, nd))
    jx = np.zeros((nx, Hp * nd))
    J = np.kron(np.eye(Hp), Dd) if nd > 0 else np.zeros((Hp * ny, 0))
    if nd > 0:
        for j in range(1, Hp + 1):
            G[ny * (j - 1):ny * j] = C @ Apow[j - 1] @ Bd
        for j in range(1, Hp + 1):
            cols = slice(nd * (j - 1), nd * j)
            J[ny * j:, cols] = G[:ny * (Hp - j)]
            jx[:, cols] = Apow[Hp - j - 1] @ Bd if j < Hp else 0.0
    f = estim.fophat - estim.xophat
    bx = S(Hp - 1) @ f
    B = np.vstack([C @ S(j) for j in range(Hp)]) @ f
    return E, G, J, K, V, B, ex, gx, jx, kx, vx, bx


@dataclasses.dataclass
class ControllerConstraint:
    """Numeric content of ``ControllerConstraint`` (src/controller/construct.jl:126-199)."""
    U0min: np.ndarray
    U0max: np.ndarray
    DUmin: np.ndarray
    DUmax: np.ndarray
    Y0min: np.ndarray
    Y0max: np.ndarray
    xhat0min: np.ndarray
    xhat0max: np.ndarray
    C_umin: np.ndarray
    C_umax: np.ndarray
    C_dumin: np.ndarray
    C_dumax: np.ndarray
    C_ymin: np.ndarray
    C_ymax: np.ndarray
    c_xmin: np.ndarray
    c_xmax: np.ndarray
    Wmin: np.ndarray = None     # custom linear constraints (nw*(Hp+1),), construct.jl:158-159
    Wmax: np.ndarray = None
    C_wmin: np.ndarray = None
    C_wmax: np.ndarray = None
    Fw: np.ndarray = None
    A: np.ndarray = None
    b: np.ndarray = None
    i_b: np.ndarray = None
    Zmin: np.ndarray = None
    Zmax: np.ndarray = None
    fx: np.ndarray = None


class LinMPC:
    """Restatement of ``LinMPC`` (src/controller/linmpc.jl:3-111, kw constructors :229-316).

    Only the SingleShooting transcription: the scope of SURVEY section 8(a), plus the custom linear
    constraints ``Wmin <= Wy Ŷe + Wu Ue + Wd D̂e + Wr R̂e <= Wmax`` of section 8(f-3) (validate_custom_lincon
    construct.jl:666-695, relaxW :1086-1160, linconstraint_custom! execute.jl:337-366), restated here ahead of the
    CUDA path and pinned by tests/test_oracle_linmpc.py to test/3_test_predictive_control.jl:466-496.
    ``estim`` is a SteadyKalmanFilter (default) or ManualEstimator.
    """

    def __init__(self, model_or_estim, Hp=None, Hc=DEFAULT_HC, Mwt=None, Nwt=None, Lwt=None,
                 M_Hp=None, N_Hc=None, L_Hp=None, Cwt=DEFAULT_CWT, Wy=None, Wu=None, Wd=None, Wr=None, **kwargs):
        estim = model_or_estim if isinstance(model_or_estim, StateEstimator) else SteadyKalmanFilter(model_or_estim, **kwargs)
        model = estim.model
        self.estim, self.model = estim, model
        nu, ny, nd, nx = model.nu, model.ny, model.nd, estim.nxhat
        if Hp is None:  # default_Hp, src/controller/construct.jl:569,581-591
            Hp = DEFAULT_HP0 + int(np.sum(np.isclose(np.abs(np.linalg.eigvals(model.A)), 0.0, atol=1e-3)))
        nb = move_blocking(Hp, Hc)
        Hc = len(nb)
        self.Hp, self.Hc, self.nb = Hp, Hc, nb
        Mwt = np.full(ny, DEFAULT_MWT) if Mwt is None else np.asarray(Mwt, float)
        Nwt = np.full(nu, DEFAULT_NWT) if Nwt is None else np.asarray(Nwt, float)
        Lwt = np.full(nu, DEFAULT_LWT) if Lwt is None else np.asarray(Lwt, float)
        for v, n, name in ((Mwt, ny, "Mwt"), (Nwt, nu, "Nwt"), (Lwt, nu, "Lwt")):  # linmpc.jl:229-316 (DimensionMismatch)
            if v.shape != (n,):
                raise ValueError(f"{name} size {v.shape} != ({n},)")
        self.M_Hp = np.diag(np.tile(Mwt, max(Hp, 0))) if M_Hp is None else np.asarray(M_Hp, float)
        self.N_Hc = np.diag(np.tile(Nwt, Hc)) if N_Hc is None else np.asarray(N_Hc, float)
        self.L_Hp = np.diag(np.tile(Lwt, max(Hp, 0))) if L_Hp is None else np.asarray(L_Hp, float)
        validate_weights(nu, ny, Hp, Hc, self.M_Hp, self.N_Hc, self.L_Hp, Cwt)
        self.Cwt = float(Cwt)
        self.neps = 0 if np.isinf(self.Cwt) else 1  # construct.jl:903
        neps = self.neps
        nZ = nu * Hc
        # ControllerWeights, construct.jl:61-93
        if neps:
            self.Ntilde_Hc = np.block([[self.N_Hc, np.zeros((nZ, 1))], [np.zeros((1, nZ)), np.array([[self.Cwt]])]])
        else:
            self.Ntilde_Hc = self.N_Hc
        PDu = init_ZtoDU(nu, Hc, nZ)
        self.Pu, self.Tu = init_ZtoU(nu, Hp, Hc, nb, nZ)
        (self.E, self.G, self.J, self.K, self.V, self.B,
         self.ex, self.gx, self.jx, self.kx, self.vx, self.bx) = init_predmat(estim, Hp, Hc, nb)
        # relax* (construct.jl:999-1199): augmented matrices for Ztilde = [Z; eps]
        if neps:
            self.Ptilde_u = np.hstack([self.Pu, np.zeros((nu * Hp, 1))])
            self.Ptilde_Du = np.block([[PDu, np.zeros((nZ, 1))], [np.zeros((1, nZ)), np.ones((1, 1))]])
            self.Etilde = np.hstack([self.E, np.zeros((ny * Hp, 1))])
            self.etilde_x = np.hstack([self.ex, np.zeros((nx, 1))])
        else:
            self.Ptilde_u, self.Ptilde_Du, self.Etilde, self.etilde_x = self.Pu, PDu, self.E, self.ex
        inf = np.inf
        # validate_custom_lincon (construct.jl:666-695): nw rows, missing matrices are zero
        given = [np.atleast_2d(np.asarray(W, float)) for W in (Wy, Wu, Wd, Wr) if W is not None]
        nw = given[0].shape[0] if given else 0

        def mat(W, ncol, name):
            if W is None:
                return np.zeros((nw, ncol))
            W = np.atleast_2d(np.asarray(W, float))
            if W.shape[1] != ncol:
                raise ValueError(f"{name} must have {ncol} columns.")
            if W.shape[0] != nw:
                raise ValueError("all custom linear constraint matrices must have the same number of rows.")
            return W
        self.Wy, self.Wu, self.Wd, self.Wr = mat(Wy, ny, "Wy"), mat(Wu, nu, "Wu"), mat(Wd, nd, "Wd"), mat(Wr, ny, "Wr")
        self.nw = nw
        rd = lambda W: np.kron(np.eye(Hp + 1), W)                    # repeatdiag(W, Hp+1), construct.jl:922-925
        self.Wbar_y, self.Wbar_u, self.Wbar_d, self.Wbar_r = rd(self.Wy), rd(self.Wu), rd(self.Wd), rd(self.Wr)
        # relaxW (construct.jl:1138-1160): Ew = W̄y [0; E] + W̄u [Pu; pu], pu = last nu rows of Pu
        self.Ew = (self.Wbar_y @ np.vstack([np.zeros((ny, nZ)), self.E])
                   + self.Wbar_u @ np.vstack([self.Pu, self.Pu[-nu:]]))
        # init_defaultcon_mpc defaults, construct.jl:904-921
        self.con = ControllerConstraint(
            U0min=np.full(nu * Hp, -inf), U0max=np.full(nu * Hp, inf),
            DUmin=np.full(nu * Hc, -inf), DUmax=np.full(nu * Hc, inf),
            Y0min=np.full(ny * Hp, -inf), Y0max=np.full(ny * Hp, inf),
            xhat0min=np.full(nx, -inf), xhat0max=np.full(nx, inf),
            C_umin=np.zeros(nu * Hp), C_umax=np.zeros(nu * Hp),
            C_dumin=np.zeros(nu * Hc), C_dumax=np.zeros(nu * Hc),
            C_ymin=np.ones(ny * Hp), C_ymax=np.ones(ny * Hp),
            c_xmin=np.ones(nx), c_xmax=np.ones(nx),
            Wmin=np.full(nw * (Hp + 1), -inf), Wmax=np.full(nw * (Hp + 1), inf),
            C_wmin=np.ones(nw * (Hp + 1)), C_wmax=np.ones(nw * (Hp + 1)), Fw=np.zeros(nw * (Hp + 1)))
        self.Uop, self.Yop, self.Dop = np.tile(model.uop, Hp), np.tile(model.yop, Hp), np.tile(model.dop, Hp)
        self._rebuild_constraints()
        # init_quadprog, construct.jl:837-845
        self.Htilde = 2 * (self.Etilde.T @ self.M_Hp @ self.Etilde
                           + self.Ptilde_Du.T @ self.Ntilde_Hc @ self.Ptilde_Du
                           + self.Ptilde_u.T @ self.L_Hp @ self.Ptilde_u)
        n = nZ + neps
        self.Ztilde = np.zeros(n)
        self.lastu0 = np.zeros(nu)
        self.qtilde, self.r = np.zeros(n), 0.0
        self.F = np.zeros(ny * Hp)
        self.d0, self.Dhat0 = np.zeros(nd), np.zeros(nd * Hp)
        self.Rhat_y, self.Rhat_u = np.zeros(ny * Hp), np.zeros(nu * Hp)
        self.Tu_lastu0 = np.zeros(nu * Hp)
        self.solved_once = False
        self.last_status = None
        self.last_qp = None

    # ---- constraints --------------------------------------------------------------
    def _rebuild_constraints(self):
        """relaxU/DU/Yhat/terminal (construct.jl:999-1199), init_boxconstraint_mpc (:1209-1234),
        init_matconstraint_mpc + deleteDU_lincon! (controller/transcription.jl:667-703,783-789)."""
        c, neps = self.con, self.neps
        nZ = self.model.nu * self.Hc
        col = lambda v: v.reshape(-1, 1)
        PDu = np.eye(nZ)
        if neps:
            A_Umin, A_Umax = -np.hstack([self.Pu, col(c.C_umin)]), np.hstack([self.Pu, -col(c.C_umax)])
            A_DUmin, A_DUmax = -np.hstack([PDu, col(c.C_dumin)]), np.hstack([PDu, -col(c.C_dumax)])
            A_Ymin, A_Ymax = -np.hstack([self.E, col(c.C_ymin)]), np.hstack([self.E, -col(c.C_ymax)])
            A_xmin, A_xmax = -np.hstack([self.ex, col(c.c_xmin)]), np.hstack([self.ex, -col(c.c_xmax)])
            A_Wmin, A_Wmax = -np.hstack([self.Ew, col(c.C_wmin)]), np.hstack([self.Ew, -col(c.C_wmax)])
        else:
            A_Wmin, A_Wmax = -self.Ew, self.Ew
            A_Umin, A_Umax = -self.Pu, self.Pu
            A_DUmin, A_DUmax = -PDu, PDu
            A_Ymin, A_Ymax = -self.E, self.E
            A_xmin, A_xmax = -self.ex, self.ex
        n = nZ + neps
        Zmin, Zmax = np.full(n, -np.inf), np.full(n, np.inf)
        if neps:
            Zmin[-1] = 0.0
            hard_min, hard_max = A_DUmin[:, -1] == 0, A_DUmax[:, -1] == 0
            Zmin[:nZ][hard_min] = c.DUmin[hard_min]
            Zmax[:nZ][hard_max] = c.DUmax[hard_max]
        else:
            Zmin[:nZ], Zmax[:nZ] = c.DUmin, c.DUmax
        fin = np.isfinite
        i_DUmin = fin(c.DUmin) & ~fin(Zmin[:nZ])
        i_DUmax = fin(c.DUmax) & ~fin(Zmax[:nZ])
        c.i_b = np.concatenate([fin(c.U0min), fin(c.U0max), i_DUmin, i_DUmax,
                                fin(c.Y0min), fin(c.Y0max), fin(c.Wmin), fin(c.Wmax),
                                fin(c.xhat0min), fin(c.xhat0max)])
        c.A = np.vstack([A_Umin, A_Umax, A_DUmin, A_DUmax, A_Ymin, A_Ymax, A_Wmin, A_Wmax, A_xmin, A_xmax])
        c.Zmin, c.Zmax = Zmin, Zmax

    def setconstraint(self, umin=None, umax=None, dumin=None, dumax=None, ymin=None, ymax=None,
                      xhatmin=None, xhatmax=None, Umin=None, Umax=None, DUmin=None, DUmax=None,
                      Ymin=None, Ymax=None, c_umin=None, c_umax=None, c_dumin=None, c_dumax=None,
                      c_ymin=None, c_ymax=None, c_xhatmin=None, c_xhatmax=None,
                      C_umin=None, C_umax=None, C_dumin=None, C_dumax=None, C_ymin=None, C_ymax=None,
                      wmin=None, wmax=None, Wmin=None, Wmax=None, c_wmin=None, c_wmax=None, C_wmin=None, C_wmax=None):
        """src/controller/construct.jl:324-559 (same argument meaning; ASCII keyword names)."""
        c, Hp, Hc = self.con, self.Hp, self.Hc
        nu, ny, nx = self.model.nu, self.model.ny, self.estim.nxhat
        old_ib, old_zmin, old_zmax = c.i_b.copy(), c.Zmin.copy(), c.Zmax.copy()
        a = lambda v: np.asarray(v, float).reshape(-1)

        def chk(v, n, name):
            v = a(v)
            if v.size != n:
                raise ValueError(f"{name} size must be ({n},)")
            return v
        if Umin is None and umin is not None: c.U0min = np.tile(chk(umin, nu, "umin"), Hp) - self.Uop
        elif Umin is not None: c.U0min = chk(Umin, nu * Hp, "Umin") - self.Uop
        if Umax is None and umax is not None: c.U0max = np.tile(chk(umax, nu, "umax"), Hp) - self.Uop
        elif Umax is not None: c.U0max = chk(Umax, nu * Hp, "Umax") - self.Uop
        if DUmin is None and dumin is not None: c.DUmin = np.tile(chk(dumin, nu, "dumin"), Hc)
        elif DUmin is not None: c.DUmin = chk(DUmin, nu * Hc, "DUmin").copy()
        if DUmax is None and dumax is not None: c.DUmax = np.tile(chk(dumax, nu, "dumax"), Hc)
        elif DUmax is not None: c.DUmax = chk(DUmax, nu * Hc, "DUmax").copy()
        if Ymin is None and ymin is not None: c.Y0min = np.tile(chk(ymin, ny, "ymin"), Hp) - self.Yop
        elif Ymin is not None: c.Y0min = chk(Ymin, ny * Hp, "Ymin") - self.Yop
        if Ymax is None and ymax is not None: c.Y0max = np.tile(chk(ymax, ny, "ymax"), Hp) - self.Yop
        elif Ymax is not None: c.Y0max = chk(Ymax, ny * Hp, "Ymax") - self.Yop
        nw = self.nw
        if Wmin is None and wmin is not None: c.Wmin = np.tile(chk(wmin, nw, "wmin"), Hp + 1)      # construct.jl:410-418
        elif Wmin is not None: c.Wmin = chk(Wmin, nw * (Hp + 1), "Wmin").copy()
        if Wmax is None and wmax is not None: c.Wmax = np.tile(chk(wmax, nw, "wmax"), Hp + 1)
        elif Wmax is not None: c.Wmax = chk(Wmax, nw * (Hp + 1), "Wmax").copy()
        if xhatmin is not None: c.xhat0min = chk(xhatmin, nx, "xhatmin") - self.estim.xophat
        if xhatmax is not None: c.xhat0max = chk(xhatmax, nx, "xhatmax") - self.estim.xophat
        ecrs = [c_umin, c_umax, c_dumin, c_dumax, c_ymin, c_ymax, c_xhatmin, c_xhatmax,
                C_umin, C_umax, C_dumin, C_dumax, C_ymin, C_ymax, c_wmin, c_wmax, C_wmin, C_wmax]
        if any(e is not None for e in ecrs):
            if self.neps != 1:
                raise ValueError("Slack variable weight Cwt must be finite to set softness parameters")
            if self.solved_once:
                raise RuntimeError("Cannot set softness parameters after calling moveinput!")
        if not self.solved_once:
            # sizes are checked before anything is stored (DimensionMismatch, construct.jl:440-506)
            rep = lambda small, big, n1, k, nm: (np.tile(chk(small, n1, nm.lower()), k) if (big is None and small is not None)
                                                 else (None if big is None else chk(big, n1 * k, nm)))
            for name, small, big, n1, k in (("C_umin", c_umin, C_umin, nu, Hp), ("C_umax", c_umax, C_umax, nu, Hp),
                                            ("C_dumin", c_dumin, C_dumin, nu, Hc), ("C_dumax", c_dumax, C_dumax, nu, Hc),
                                            ("C_ymin", c_ymin, C_ymin, ny, Hp), ("C_ymax", c_ymax, C_ymax, ny, Hp),
                                            ("C_wmin", c_wmin, C_wmin, nw, Hp + 1), ("C_wmax", c_wmax, C_wmax, nw, Hp + 1)):
                v = rep(small, big, n1, k, name)
                if v is not None:
                    if (v < 0).any():
                        raise ValueError(f"{name} weights should be non-negative")
                    setattr(c, name, v)
            if c_xhatmin is not None: c.c_xmin = chk(c_xhatmin, nx, "c_xhatmin")
            if c_xhatmax is not None: c.c_xmax = chk(c_xhatmax, nx, "c_xhatmax")
        self._rebuild_constraints()
        if self.solved_once:
            # construct.jl:548-551: the +-Inf pattern is frozen after the first solve
            if ((c.i_b != old_ib).any() or (np.isinf(c.Zmin) != np.isinf(old_zmin)).any()
                    or (np.isinf(c.Zmax) != np.isinf(old_zmax)).any()):
                raise RuntimeError("Cannot modify +-Inf constraints after calling moveinput!")
        return self

    # ---- per-step path --------------------------------------------------------------
    def initpred(self, ry, d, lastu, Dhat, Rhat_y, Rhat_u):
        """initpred_common! + initpred! (src/controller/execute.jl:297-314, 247-277)."""
        m = self.model
        self.lastu0 = np.asarray(lastu, float) - m.uop
        self.Tu_lastu0 = self.Tu @ self.lastu0
        self.yhat = self.estim.evaloutput(d)
        if m.nd > 0:
            self.d0 = np.asarray(d, float) - m.dop
            self.Dhat0 = np.asarray(Dhat, float) - self.Dop
        self.Rhat_y, self.Rhat_u = np.asarray(Rhat_y, float), np.asarray(Rhat_u, float)
        self.ry = np.asarray(ry, float).reshape(-1)
        # F starts from the stochastic predictions Ŷs of an InternalModel (predictstoch!, execute.jl:321-327), else from 0
        Ks, Ps = init_stochpred(self.estim, self.Hp)
        F = Ks @ self.estim.xs + Ps @ self.estim.ys if Ks.shape[0] else 0.0
        F = F + self.B + self.K @ self.estim.xhat0 + self.V @ self.lastu0
        if m.nd > 0:
            F = F + self.G @ self.d0 + self.J @ self.Dhat0
        self.F = F
        q = np.zeros_like(self.qtilde)
        r = 0.0
        if np.any(self.M_Hp):
            Cy = F + self.Yop - self.Rhat_y
            q = q + (self.M_Hp @ self.Etilde).T @ Cy
            r += Cy @ self.M_Hp @ Cy
        if np.any(self.L_Hp):
            Cu = self.Tu_lastu0 + self.Uop - self.Rhat_u
            q = q + (self.L_Hp @ self.Ptilde_u).T @ Cu
            r += Cu @ self.L_Hp @ Cu
        self.qtilde, self.r = 2 * q, r

    def linconstraint(self):
        """src/controller/transcription.jl:811-848."""
        c = self.con
        fx = self.bx + self.kx @ self.estim.xhat0 + self.vx @ self.lastu0
        if self.model.nd > 0:
            fx = fx + self.gx @ self.d0 + self.jx @ self.Dhat0
        c.fx = fx
        # linconstraint_custom! + linconstraint_custom_outputs! (execute.jl:337-366): Fw in ABSOLUTE units
        m = self.model
        Fw = np.zeros(self.nw * (self.Hp + 1))
        if self.nw:
            Ue = np.concatenate([self.Tu_lastu0 + self.Uop, self.lastu0 + m.uop])
            Fw = Fw + self.Wbar_u @ Ue
            if m.nd > 0:
                Fw = Fw + self.Wbar_d @ np.concatenate([self.d0 + m.dop, self.Dhat0 + self.Dop])
            Fw = Fw + self.Wbar_r @ np.concatenate([self.ry, self.Rhat_y])
            Fw = Fw + self.Wbar_y @ np.concatenate([self.yhat, self.F + self.Yop])
        c.Fw = Fw
        c.b = np.concatenate([-c.U0min + self.Tu_lastu0, c.U0max - self.Tu_lastu0,
                              -c.DUmin, c.DUmax,
                              -c.Y0min + self.F, c.Y0max - self.F,
                              -c.Wmin + Fw, c.Wmax - Fw,
                              -c.xhat0min + fx, c.xhat0max - fx])

    def warmstart(self):
        """set_warmstart_mpc!, src/controller/transcription.jl:997-1007."""
        nu, nDU = self.model.nu, self.model.nu * self.Hc
        Zs = np.zeros_like(self.Ztilde)
        Zs[:nDU - nu] = self.Ztilde[nu:nDU]
        if self.neps:
            Zs[-1] = self.Ztilde[-1]
        return Zs

    def optim_objective(self):
        """src/controller/execute.jl:466-505 with the exact QP solver in place of JuMP+OSQP."""
        c = self.con
        Zs = self.warmstart()
        sol = _qp.solve_qp(self.Htilde, self.qtilde, c.A[c.i_b], c.b[c.i_b], c.Zmin, c.Zmax)
        self.last_qp, self.last_status = sol, sol["status"]
        self.solved_once = True
        if sol["status"] == _qp.INFEASIBLE:  # iserror -> shifted last solution (:499-500)
            self.Ztilde = Zs
        else:
            self.Ztilde = sol["z"].copy()
        return self.Ztilde

    def moveinput(self, ry=None, d=(), lastu=None, Dhat=None, Rhat_y=None, Rhat_u=None):
        """src/controller/execute.jl:59-80."""
        m = self.model
        ry = m.yop if ry is None else np.asarray(ry, float).reshape(-1)
        d = np.asarray(d, float).reshape(-1)
        lastu = self.lastu0 + m.uop if lastu is None else np.asarray(lastu, float).reshape(-1)
        Dhat = np.tile(d, self.Hp) if Dhat is None else np.asarray(Dhat, float)
        Rhat_y = np.tile(ry, self.Hp) if Rhat_y is None else np.asarray(Rhat_y, float)
        Rhat_u = self.Uop if Rhat_u is None else np.asarray(Rhat_u, float)
        # validate_args, construct.jl:702-710
        if ry.size != m.ny: raise ValueError("ry size")
        if d.size != m.nd: raise ValueError("d size")
        if lastu.size != m.nu: raise ValueError("lastu size")
        if Dhat.size != m.nd * self.Hp: raise ValueError("Dhat size")
        if Rhat_y.size != m.ny * self.Hp: raise ValueError("Rhaty size")
        if Rhat_u.size != m.nu * self.Hp: raise ValueError("Rhatu size")
        self.initpred(ry, d, lastu, Dhat, Rhat_y, Rhat_u)
        self.linconstraint()
        Z = self.optim_objective()
        # getinput!, execute.jl:536-546
        u = Z[:m.nu] + self.lastu0 + m.uop
        self.lastu0 = u - m.uop
        return u

    def getinfo(self):
        """src/controller/execute.jl:145-198 (+ predict! transcription.jl:1136-1145,
        obj_nonlinprog! execute.jl:415-446).  Note ``lastu0`` was already advanced by getinput!,
        but ``Tu_lastu0`` still holds the value used in the optimisation, as in the reference."""
        m, Z = self.model, self.Ztilde
        U0 = self.Ptilde_u @ Z + self.Tu_lastu0
        Y0 = self.Etilde @ Z + self.F
        xend = self.etilde_x @ Z + self.con.fx
        Ybar = Y0 + self.Yop - self.Rhat_y
        Ubar = U0 + self.Uop - self.Rhat_u
        J = Ybar @ self.M_Hp @ Ybar + Z @ self.Ntilde_Hc @ Z + Ubar @ self.L_Hp @ Ubar
        nDU = m.nu * self.Hc
        return dict(DU=Z[:nDU].copy(), eps=(Z[-1] if self.neps else 0.0), J=J, U=U0 + self.Uop,
                    u=(U0 + self.Uop)[:m.nu], Yhat=Y0 + self.Yop, xhatend=xend + self.estim.xophat,
                    yhat=self.yhat, Rhat_y=self.Rhat_y, Rhat_u=self.Rhat_u,
                    J_quad=0.5 * Z @ self.Htilde @ Z + self.qtilde @ Z + self.r)

    # estimator pass-throughs (src/controller/execute.jl:523-555)
    def preparestate(self, ym, d=()):
        return self.estim.preparestate(ym, d)

    def updatestate(self, u, ym, d=()):
        return self.estim.updatestate(u, ym, d)

    def setstate(self, xhat, *cov):
        self.estim.setstate(xhat, *cov)
        return self

    def initstate(self, u, ym, d=()):
        """initstate!(mpc, u, ym, d) (src/controller/execute.jl:11-23): the estimator's steady state, and the warm start of
        the decision vector cleared."""
        self.Ztilde[:] = 0.0
        self.lastu0 = np.asarray(u, float).reshape(-1) - self.model.uop
        return self.estim.initstate(u, ym, d)


def _setmodel_linmpc(self, model=None, Mwt=None, Nwt=None, Lwt=None, M_Hp=None, Ntilde_Hc=None, L_Hp=None, **kw):
    """setmodel!(mpc, model; Mwt, Nwt, Lwt, M_Hp, Ñ_Hc, L_Hp) and setmodel_controller! (src/controller/execute.jl:621-790):
    new plant model / weights at run time; Z̃ is kept, u0(k-1) and the deviation-form bounds are re-expressed around the
    new operating points, the prediction matrices, the constraint matrices and H̃ are rebuilt."""
    m, estim = self.model, self.estim
    nu, ny, Hp, Hc, neps = m.nu, m.ny, self.Hp, self.Hc, self.neps
    uop_old, yop_old, xop_old = m.uop.copy(), m.yop.copy(), estim.xophat.copy()
    estim.setmodel(m if model is None else model, **kw)
    m = estim.model
    diag = lambda w, n, reps, name: np.diag(np.tile(_chk_w(w, n, name), reps))
    if M_Hp is None and Mwt is not None:
        self.M_Hp = diag(Mwt, ny, Hp, "Mwt")
    elif M_Hp is not None:
        self.M_Hp = _herm(M_Hp, ny * Hp, "M_Hp")
    if Ntilde_Hc is None and Nwt is not None:
        self.Ntilde_Hc[:nu * Hc, :nu * Hc] = diag(Nwt, nu, Hc, "Nwt")
    elif Ntilde_Hc is not None:
        self.Ntilde_Hc = _herm(Ntilde_Hc, nu * Hc + neps, "Ñ_Hc")
    self.N_Hc = self.Ntilde_Hc[:nu * Hc, :nu * Hc]
    if L_Hp is None and Lwt is not None:
        self.L_Hp = diag(Lwt, nu, Hp, "Lwt")
    elif L_Hp is not None:
        self.L_Hp = _herm(L_Hp, nu * Hp, "L_Hp")
    # ---- setmodel_controller! ----
    nZ = nu * Hc
    (self.E, self.G, self.J, self.K, self.V, self.B,
     self.ex, self.gx, self.jx, self.kx, self.vx, self.bx) = init_predmat(estim, Hp, Hc, self.nb)
    if neps:
        self.Etilde = np.hstack([self.E, np.zeros((ny * Hp, 1))])
        self.etilde_x = np.hstack([self.ex, np.zeros((estim.nxhat, 1))])
    else:
        self.Etilde, self.etilde_x = self.E, self.ex
    self.Ew = (self.Wbar_y @ np.vstack([np.zeros((ny, nZ)), self.E]) + self.Wbar_u @ np.vstack([self.Pu, self.Pu[-nu:]]))
    c = self.con
    c.U0min, c.U0max = c.U0min + self.Uop, c.U0max + self.Uop            # to absolute with the OLD operating points
    c.Y0min, c.Y0max = c.Y0min + self.Yop, c.Y0max + self.Yop
    c.xhat0min, c.xhat0max = c.xhat0min + xop_old, c.xhat0max + xop_old
    self.lastu0 = self.lastu0 + uop_old - m.uop
    self.Uop, self.Yop, self.Dop = np.tile(m.uop, Hp), np.tile(m.yop, Hp), np.tile(m.dop, Hp)
    c.U0min, c.U0max = c.U0min - self.Uop, c.U0max - self.Uop            # back to deviations with the NEW ones
    c.Y0min, c.Y0max = c.Y0min - self.Yop, c.Y0max - self.Yop
    c.xhat0min, c.xhat0max = c.xhat0min - estim.xophat, c.xhat0max - estim.xophat
    self._rebuild_constraints()
    self.Htilde = 2 * (self.Etilde.T @ self.M_Hp @ self.Etilde + self.Ptilde_Du.T @ self.Ntilde_Hc @ self.Ptilde_Du
                       + self.Ptilde_u.T @ self.L_Hp @ self.Ptilde_u)
    return self


def _chk_w(w, n, name):
    w = np.asarray(w, float).reshape(-1)
    if w.shape != (n,):
        raise ValueError(f"{name} should be a vector of length {n}")
    if (w < 0).any():
        raise ValueError(f"{name} values should be nonnegative")
    return w


def _herm(M, n, name):
    M = np.asarray(M, float)
    if M.ndim == 1:
        M = np.diag(M)
    if M.shape != (n, n):
        raise ValueError(f"{name} size should be ({n}, {n})")
    return np.tril(M) + np.tril(M, -1).T


LinMPC.setmodel = _setmodel_linmpc


class ExplicitMPC(LinMPC):
    """src/controller/explicitmpc.jl:63,209: Ztilde = -H^-1 q (no constraints, Cwt = Inf)."""

    def __init__(self, model_or_estim, **kw):
        kw["Cwt"] = np.inf
        super().__init__(model_or_estim, **kw)

    def optim_objective(self):
        self.Ztilde = -np.linalg.solve(self.Htilde, self.qtilde)
        self.solved_once = True
        return self.Ztilde

    def setconstraint(self, **kw):
        """src/controller/explicitmpc.jl:181."""
        raise RuntimeError("ExplicitMPC does not support constraints.")
