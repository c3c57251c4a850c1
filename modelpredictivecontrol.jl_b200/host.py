"""Host-side helpers that mirror the reference's constructors (one-off, not on the hot path)."""
import numpy as np


def move_blocking(Hp, Hc):
    """move_blocking (reference src/controller/construct.jl:629-660)."""
    if np.isscalar(Hc):
        Hc = int(Hc)
        if Hc < 1:
            raise ValueError("Control horizon Hc should be >= 1")
        nb = [1] * Hc
        nb[-1] = Hp - Hc + 1
        if nb[-1] < 1:
            raise ValueError("Control horizon Hc should be <= prediction horizon Hp")
        return nb
    nb = [int(v) for v in Hc]
    if not all(v > 0 for v in nb):
        raise ValueError("Move blocking vector must be strictly positive integers.")
    if sum(nb) < Hp:
        nb = nb + [Hp - sum(nb)]
    elif sum(nb) > Hp:
        cs = np.cumsum(nb)
        nb = nb[: int(np.flatnonzero(cs >= Hp)[0]) + 1]
        if sum(nb) > Hp:
            nb[-1] = Hp - sum(nb[:-1])
    return nb


# ----------------------------------------------------------------------------------------------
# Batched host-side constructors (numpy, leading axis = instance).  They mirror the reference's
# one-off setup code; nothing here runs per control period except the tiny estimator updates.
# ----------------------------------------------------------------------------------------------
def _b(a, N, shape, strict=False):
    """Broadcast a per-model array to (N, *shape).  ``strict``: a shared array must have exactly ``shape`` (no numpy
    broadcasting of length-1 axes: the reference's DimensionMismatch for `setconstraint!` arguments)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == len(shape):
        if strict and a.shape != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
        a = np.broadcast_to(a, (N,) + tuple(shape))
    if a.shape != (N,) + tuple(shape):
        raise ValueError(f"expected shape {(N,) + tuple(shape)}, got {a.shape}")
    return np.ascontiguousarray(a)


def expand_softness(small, big, n, reps, name):
    """Softness (ECR) weights of one constraint family over the horizon: the per-sample form ``c_xxx`` (length ``n``,
    repeated ``reps`` times) or the whole-horizon form ``C_xxx`` (length ``n * reps``), which wins when both are given
    (``setconstraint!``, src/controller/construct.jl:440-506).  Returns None when neither is given; raises ValueError for a
    wrong size (the reference's DimensionMismatch) or a negative weight."""
    if big is not None:
        v = np.asarray(big, dtype=np.float64).reshape(-1)
        if v.size != n * reps:
            raise ValueError(f"C_{name} size must be ({n * reps},)")
    elif small is not None:
        v = np.asarray(small, dtype=np.float64).reshape(-1)
        if v.size != n:
            raise ValueError(f"c_{name} size must be ({n},)")
        v = np.tile(v, reps)
    else:
        return None
    if (v < 0).any():
        raise ValueError(f"C_{name} weights should be non-negative")
    return v.copy()


class LinModel:
    """Batch of N linear plants ``x0(k+1) = A x0 + Bu u0 + Bd d0 + fop - xop``, ``y0 = C x0 + Dd d0``
    (reference src/model/linmodel.jl:1-66, direct-matrix constructor :252-253; operating points
    as set by ``setop!`` src/sim_model.jl:101-126).  Matrices are (N, rows, cols) or 2-D (shared)."""

    def __init__(self, A, Bu, C, Bd=None, Dd=None, Ts=1.0, N=None, uop=None, yop=None, dop=None, xop=None, fop=None):
        A = np.asarray(A, dtype=np.float64)
        if N is None:
            N = A.shape[0] if A.ndim == 3 else 1
        self.N = N
        nx = A.shape[-1]
        Bu = np.asarray(Bu, dtype=np.float64)
        C = np.asarray(C, dtype=np.float64)
        nu, ny = Bu.shape[-1], C.shape[-2]
        nd = 0 if Bd is None else np.asarray(Bd).shape[-1]
        self.nx, self.nu, self.ny, self.nd, self.Ts = nx, nu, ny, nd, Ts
        self.A, self.Bu, self.C = _b(A, N, (nx, nx)), _b(Bu, N, (nx, nu)), _b(C, N, (ny, nx))
        self.Bd = _b(np.zeros((nx, 0)) if Bd is None else Bd, N, (nx, nd))
        self.Dd = _b(np.zeros((ny, nd)) if Dd is None else Dd, N, (ny, nd))
        z = lambda v, n: _b(np.zeros(n) if v is None else v, N, (n,))
        self.uop, self.yop, self.dop, self.xop, self.fop = z(uop, nu), z(yop, ny), z(dop, nd), z(xop, nx), z(fop, nx)
        self.x0 = np.zeros((N, nx))

    def setstate(self, x):
        self.x0 = _b(x, self.N, (self.nx,)) - self.xop

    def evaloutput(self, d=None):
        d0 = (np.zeros((self.N, 0)) if d is None else _b(d, self.N, (self.nd,))) - self.dop
        return np.einsum("nij,nj->ni", self.C, self.x0) + np.einsum("nij,nj->ni", self.Dd, d0) + self.yop

    def updatestate(self, u, d=None):
        u0 = _b(u, self.N, (self.nu,)) - self.uop
        d0 = (np.zeros((self.N, 0)) if d is None else _b(d, self.N, (self.nd,))) - self.dop
        self.x0 = (np.einsum("nij,nj->ni", self.A, self.x0) + np.einsum("nij,nj->ni", self.Bu, u0)
                   + np.einsum("nij,nj->ni", self.Bd, d0) + self.fop - self.xop)
        return self.x0 + self.xop


def init_integrators(nint, ny):
    """reference src/estimator/construct.jl:226-251."""
    nint = np.zeros(ny, int) if np.isscalar(nint) and nint == 0 else np.asarray(nint, int).reshape(-1)
    if nint.size != ny or (nint < 0).any():
        raise ValueError("nint must have one non-negative entry per output")
    nxs = int(nint.sum())
    A, C = np.zeros((nxs, nxs)), np.zeros((ny, nxs))
    i0 = 0
    for i, k in enumerate(nint):
        if k:
            A[i0:i0 + k, i0:i0 + k] = np.eye(k) + np.eye(k, k=-1)
            C[i, i0 + k - 1] = 1.0
            i0 += k
    return A, C, nint


def augment_model(model, nint_u=0, nint_ym=None, i_ym=None):
    """init_estimstoch + augment_model (reference src/estimator/construct.jl:172-185, 305-323), batched.
    The default ``nint_ym`` is one integrator per measured output (default_nint :365-376 additionally drops
    integrators that break observability; that check is left to the caller for synthetic plants)."""
    N, nx, nu, ny, nd = model.N, model.nx, model.nu, model.ny, model.nd
    i_ym = list(range(ny)) if i_ym is None else list(i_ym)
    if nint_ym is None:
        nint_ym = [1] * len(i_ym)
    As_u, Cs_u, _ = init_integrators(nint_u, nu)
    As_ym, Cs_ym, _ = init_integrators(nint_ym, len(i_ym))
    Cs_y = np.zeros((ny, Cs_ym.shape[1]))
    Cs_y[i_ym] = Cs_ym
    nsu, nsy = As_u.shape[0], As_ym.shape[0]
    nxs = nsu + nsy
    As = np.zeros((nxs, nxs))
    As[:nsu, :nsu], As[nsu:, nsu:] = As_u, As_ym
    Cs_u = np.hstack([Cs_u, np.zeros((nu, nsy))])
    Cs_y = np.hstack([np.zeros((ny, nsu)), Cs_y])
    nxh = nx + nxs
    Ahat = np.zeros((N, nxh, nxh))
    Ahat[:, :nx, :nx] = model.A
    Ahat[:, :nx, nx:] = model.Bu @ Cs_u
    Ahat[:, nx:, nx:] = As
    Buhat = np.concatenate([model.Bu, np.zeros((N, nxs, nu))], axis=1)
    Chat = np.concatenate([model.C, np.broadcast_to(Cs_y, (N, ny, nxs))], axis=2)
    Bdhat = np.concatenate([model.Bd, np.zeros((N, nxs, nd))], axis=1)
    xop = np.concatenate([model.xop, np.zeros((N, nxs))], axis=1)
    fop = np.concatenate([model.fop, np.zeros((N, nxs))], axis=1)
    return dict(Ahat=Ahat, Buhat=Buhat, Chat=Chat, Bdhat=Bdhat, Ddhat=model.Dd.copy(), xophat=xop, fophat=fop,
                nxhat=nxh, nxs=nxs, nsu=nsu, i_ym=i_ym, nint_ym=list(nint_ym), nint_u=nint_u)


def dare_filter_sda(A, C, Q, R, iters=60, tol=1e-13):
    """Batched filter Riccati  P = A P A' - A P C'(C P C' + R)^-1 C P A' + Q  by the structure-preserving
    doubling algorithm (quadratic convergence); A (N,n,n), C (N,m,n), Q (n,n) or (N,n,n), R (m,m) or (N,m,m).
    Stands in for ControlSystemsBase.kalman (third-party) used by init_skf, reference
    src/estimator/kalman.jl:204-227."""
    N, n = A.shape[0], A.shape[-1]
    Q = np.broadcast_to(Q, (N, n, n))
    R = np.broadcast_to(R, (N,) + R.shape[-2:])
    Ak = np.swapaxes(A, 1, 2).copy()                       # dual system: A -> A', B -> C'
    Gk = np.swapaxes(C, 1, 2) @ np.linalg.solve(R, C)       # C' R^-1 C
    Hk = Q.copy()
    I = np.eye(n)
    for _ in range(iters):
        W = np.linalg.inv(I + Gk @ Hk)
        AW = Ak @ W
        Gn = Gk + AW @ Gk @ np.swapaxes(Ak, 1, 2)
        Hn = Hk + np.swapaxes(Ak, 1, 2) @ Hk @ W @ Ak
        An = AW @ Ak
        done = np.abs(Hn - Hk).max() <= tol * (1 + np.abs(Hn).max())
        Ak, Gk, Hk = An, Gn, Hn
        if done:
            break
    return 0.5 * (Hk + np.swapaxes(Hk, 1, 2))


class SteadyKalmanFilter:
    """Batched ``SteadyKalmanFilter`` (reference src/estimator/kalman.jl:163-227, 284-309): constant gain
    K̂ = P Ĉm'(Ĉm P Ĉm' + R̂)^-1 (direct=true).  ``preparestate`` / ``updatestate`` are the reference's
    correct_estimate_obsv! / predict_estimate_obsv!."""

    direct = True

    def __init__(self, model, nint_u=0, nint_ym=None, i_ym=None, sigmaQ=None, sigmaR=None, sigmaQint_u=None,
                 sigmaQint_ym=None):
        self.model = model
        aug = augment_model(model, nint_u, nint_ym, i_ym)
        self.__dict__.update(aug)
        N, nx = model.N, model.nx
        nym = len(self.i_ym)
        sQ = np.full(nx, 1.0 / nx) if sigmaQ is None else np.asarray(sigmaQ, float)
        sR = np.ones(nym) if sigmaR is None else np.asarray(sigmaR, float)
        nsu = self.nsu  # integrator states on the manipulated inputs (init_integrators)
        sQu = np.ones(nsu) if sigmaQint_u is None else np.asarray(sigmaQint_u, float)
        sQy = np.ones(self.nxs - nsu) if sigmaQint_ym is None else np.asarray(sigmaQint_ym, float)
        self.Qhat = np.diag(np.concatenate([sQ, sQu, sQy]) ** 2)
        self.Rhat = np.diag(sR ** 2)
        self.Cmhat, self.Ddmhat = self.Chat[:, self.i_ym], self.Ddhat[:, self.i_ym]
        P = dare_filter_sda(self.Ahat, self.Cmhat, self.Qhat, self.Rhat)
        S = self.Cmhat @ P @ np.swapaxes(self.Cmhat, 1, 2) + self.Rhat
        self.Khat = np.swapaxes(np.linalg.solve(np.swapaxes(S, 1, 2), self.Cmhat @ np.swapaxes(P, 1, 2)), 1, 2)
        self.Phat = P
        self.xhat0 = np.zeros((N, self.nxhat))

    def _d0(self, d):
        return (np.zeros((self.model.N, 0)) if d is None else _b(d, self.model.N, (self.model.nd,))) - self.model.dop

    def setstate(self, xhat):
        self.xhat0 = _b(xhat, self.model.N, (self.nxhat,)) - self.xophat

    def setmodel(self, model):
        """setmodel!(estim, model) (reference src/estimator/execute.jl:483-544), batched: new plant matrices and operating
        points, same stochastic model; x̂0 is re-expressed around the new x̂op.  SteadyKalmanFilter itself refuses, as the
        reference does (kalman.jl:229-234); KalmanFilter and ManualEstimator accept."""
        if type(self) is SteadyKalmanFilter:
            raise RuntimeError("SteadyKalmanFilter does not support setmodel! (use KalmanFilter instead)")
        old = self.model
        if (model.N, model.nx, model.nu, model.ny, model.nd) != (old.N, old.nx, old.nu, old.ny, old.nd):
            raise ValueError("model dimensions must be the same")
        xhat = self.xhat0 + self.xophat
        keep = {k: getattr(self, k) for k in ("xhat0",)}
        self.model = model
        self.__dict__.update(augment_model(model, self.nint_u, self.nint_ym, self.i_ym))
        self.Cmhat, self.Ddmhat = self.Chat[:, self.i_ym], self.Ddhat[:, self.i_ym]
        self.xhat0 = xhat - self.xophat
        return self

    def preparestate(self, ym, d=None):
        y0m = _b(ym, self.model.N, (len(self.i_ym),)) - self.model.yop[:, self.i_ym]
        d0 = self._d0(d)
        v = y0m - (np.einsum("nij,nj->ni", self.Cmhat, self.xhat0) + np.einsum("nij,nj->ni", self.Ddmhat, d0))
        # an instance with a NaN measurement skips its correction step (kalman.jl:245-251, :478-484)
        self._nan_ym = np.isnan(y0m).any(axis=1)
        if self._nan_ym.any():
            v = np.where(self._nan_ym[:, None], 0.0, v)
        self.xhat0 = self.xhat0 + np.einsum("nij,nj->ni", self.Khat, v)
        return self.xhat0 + self.xophat

    def initstate(self, u, ym, d=None):
        """initstate! (init_estimate!, src/estimator/execute.jl:246-259): the steady state of the augmented model for the
        inputs u, d that reproduces the measurement ym (least squares, one system per instance)."""
        m, N = self.model, self.model.N
        u0 = _b(u, N, (m.nu,)) - m.uop
        y0m = _b(ym, N, (len(self.i_ym),)) - m.yop[:, self.i_ym]
        d0 = self._d0(d)
        rhs_x = (self.fophat - self.xophat + np.einsum("nij,nj->ni", self.Buhat, u0) + np.einsum("nij,nj->ni", self.Bdhat, d0))
        rhs_y = y0m - np.einsum("nij,nj->ni", self.Ddmhat, d0)
        M = np.concatenate([np.eye(self.nxhat) - self.Ahat, self.Cmhat], axis=1)
        rhs = np.concatenate([rhs_x, rhs_y], axis=1)
        self.xhat0 = np.stack([np.linalg.lstsq(M[i], rhs[i], rcond=None)[0] for i in range(N)])
        return self.xhat0 + self.xophat

    def updatestate(self, u, ym, d=None):
        u0 = _b(u, self.model.N, (self.model.nu,)) - self.model.uop
        d0 = self._d0(d)
        self.xhat0 = (np.einsum("nij,nj->ni", self.Ahat, self.xhat0) + np.einsum("nij,nj->ni", self.Buhat, u0)
                      + np.einsum("nij,nj->ni", self.Bdhat, d0) + self.fophat - self.xophat)
        return self.xhat0 + self.xophat

    def evaloutput(self, d=None):
        d0 = self._d0(d)
        return (np.einsum("nij,nj->ni", self.Chat, self.xhat0) + np.einsum("nij,nj->ni", self.Ddhat, d0)
                + self.model.yop)


class KalmanFilter(SteadyKalmanFilter):
    """Batched time-varying ``KalmanFilter`` (reference src/estimator/kalman.jl:311-525, direct=true):
    ``preparestate`` = correct_estimate_kf! (:1235-1268, gain from the current P̂, P̂ <- Hermitian((I - K̂ Ĉm) P̂, :L)),
    ``updatestate`` = predict_estimate_kf! (:1270-1290, P̂ <- Hermitian(Â P̂ Â' + Q̂, :L)).  Host-side numpy version for
    the x̂0-input seam; ``LinMPC(KalmanFilter(model), fused_estimator=True)`` runs the same recursion on the GPU."""

    def __init__(self, model, nint_u=0, nint_ym=None, i_ym=None, sigmaP_0=None, sigmaQ=None, sigmaR=None,
                 sigmaPint_ym_0=None, sigmaQint_ym=None):
        self.model = model
        self.__dict__.update(augment_model(model, nint_u, nint_ym, i_ym))
        N, nx = model.N, model.nx
        nym = len(self.i_ym)
        one = lambda v, n, d: np.full(n, d) if v is None else np.asarray(v, float).reshape(n)
        sP = np.concatenate([one(sigmaP_0, nx, 1.0 / nx), one(sigmaPint_ym_0, self.nxs, 1.0)])
        sQ = np.concatenate([one(sigmaQ, nx, 1.0 / nx), one(sigmaQint_ym, self.nxs, 1.0)])
        self.P0hat, self.Qhat, self.Rhat = np.diag(sP ** 2), np.diag(sQ ** 2), np.diag(one(sigmaR, nym, 1.0) ** 2)
        self.Cmhat, self.Ddmhat = self.Chat[:, self.i_ym], self.Ddhat[:, self.i_ym]
        self.Phat = np.broadcast_to(self.P0hat, (N, self.nxhat, self.nxhat)).copy()
        self.Khat = np.zeros((N, self.nxhat, nym))
        self.xhat0 = np.zeros((N, self.nxhat))

    @staticmethod
    def _herm_lower(P):
        L = np.tril(P)
        return L + np.swapaxes(np.tril(P, -1), 1, 2)

    def preparestate(self, ym, d=None):
        Cm, P = self.Cmhat, self.Phat
        PCt = P @ np.swapaxes(Cm, 1, 2)
        M = Cm @ PCt + self.Rhat
        self.Khat = np.swapaxes(np.linalg.solve(np.swapaxes(M, 1, 2), np.swapaxes(PCt, 1, 2)), 1, 2)
        out = SteadyKalmanFilter.preparestate(self, ym, d)
        Pc = self._herm_lower((np.eye(self.nxhat) - self.Khat @ Cm) @ P)
        self.Phat = np.where(self._nan_ym[:, None, None], P, Pc) if self._nan_ym.any() else Pc  # (no correction on NaN)
        return out

    def updatestate(self, u, ym, d=None):
        out = SteadyKalmanFilter.updatestate(self, u, ym, d)
        self.Phat = self._herm_lower(self.Ahat @ self.Phat @ np.swapaxes(self.Ahat, 1, 2) + self.Qhat)
        return out


class ManualEstimator(SteadyKalmanFilter):
    """``ManualEstimator`` (reference src/estimator/manual.jl:60-64,150-154): the caller sets x̂ with
    ``setstate``; prepare/update do nothing."""

    def __init__(self, model, nint_u=0, nint_ym=None, i_ym=None):
        self.model = model
        self.__dict__.update(augment_model(model, nint_u, nint_ym, i_ym))
        self.xhat0 = np.zeros((model.N, self.nxhat))

    def preparestate(self, ym, d=None):
        return self.xhat0 + self.xophat

    def updatestate(self, u, ym, d=None):
        return self.xhat0 + self.xophat


class InternalModel(SteadyKalmanFilter):
    """Batched ``InternalModel`` estimator (reference src/estimator/internal_model.jl): no state augmentation
    (Â = A, ...), a stochastic model of the measured outputs (default one integrator per measured output,
    ``stoch_ym = (As, Bs, Cs, Ds)`` shared by the batch otherwise) updated by  x̂s <- Âs x̂s + B̂s ŷs  with
    ŷs^m = ym - ŷd^m (correct_estimate! :262-277, update_estimate! :293-311).  The controller adds the stochastic
    predictions Ŷs = Ks x̂s + Ps ŷs (init_stochpred, construct.jl:1254-1267) to F through ``io.Yhat_s``."""

    def __init__(self, model, i_ym=None, stoch_ym=None):
        if np.any(np.abs(np.linalg.eigvals(model.A)) >= 1):
            raise ValueError("InternalModel does not support integrating or unstable model")
        self.model = model
        N, ny = model.N, model.ny
        self.i_ym = list(range(ny)) if i_ym is None else list(i_ym)
        nym = len(self.i_ym)
        if stoch_ym is None:
            Asm = Bsm = Csm = Dsm = np.eye(nym)
        else:
            Asm, Bsm, Csm, Dsm = [np.atleast_2d(np.asarray(M, float)) for M in stoch_ym]
        if not np.any(Dsm):
            raise ValueError("Stochastic model requires a nonzero direct transmission matrix D")
        nxs = Asm.shape[0]
        Bs, Cs, Ds = np.zeros((nxs, ny)), np.zeros((ny, nxs)), np.eye(ny)
        Bs[:, self.i_ym] = Bsm
        Cs[self.i_ym] = Csm
        Ds[np.ix_(self.i_ym, self.i_ym)] = Dsm
        self.As, self.Cs = Asm, Cs
        self.Bs_hat = Bs @ np.linalg.inv(Ds)
        self.As_hat = Asm - self.Bs_hat @ Cs
        self.nxs, self.nxhat, self.nsu = nxs, model.nx, 0
        self.nint_u, self.nint_ym = 0, [0] * nym
        self._set_matrices()
        self.xhat0 = np.zeros((N, model.nx))
        self.xs = np.zeros((N, nxs))
        self.ys = np.zeros((N, ny))

    def _set_matrices(self):
        m = self.model
        self.Ahat, self.Buhat, self.Chat, self.Bdhat, self.Ddhat = m.A, m.Bu, m.C, m.Bd, m.Dd
        self.xophat, self.fophat = m.xop.copy(), m.fop.copy()
        self.Cmhat, self.Ddmhat = self.Chat[:, self.i_ym], self.Ddhat[:, self.i_ym]

    def setmodel(self, model):
        xhat = self.xhat0 + self.xophat
        self.model = model
        self._set_matrices()
        self.xhat0 = xhat - self.xophat
        return self

    def stochpred(self, Hp):
        """init_stochpred: (Ks, Ps) with Ŷs = Ks x̂s + Ps ŷs (shared by the batch)."""
        ny = self.model.ny
        Ks, Ps = np.zeros((ny * Hp, self.nxs)), np.zeros((ny * Hp, ny))
        Ap = np.eye(self.nxs)
        for i in range(1, Hp + 1):
            Ms = self.Cs @ Ap @ self.Bs_hat
            Ap = Ap @ self.As
            Ks[ny * (i - 1):ny * i] = self.Cs @ Ap - Ms @ self.Cs
            Ps[ny * (i - 1):ny * i] = Ms
        return Ks, Ps

    def preparestate(self, ym, d=None):
        m, N = self.model, self.model.N
        y0m = _b(ym, N, (len(self.i_ym),)) - m.yop[:, self.i_ym]
        d0 = self._d0(d)
        yd = np.einsum("nij,nj->ni", self.Chat, self.xhat0) + np.einsum("nij,nj->ni", self.Ddhat, d0)
        ys = np.zeros((N, m.ny))
        ys[:, self.i_ym] = np.where(np.isfinite(y0m), y0m - yd[:, self.i_ym], 0.0)
        self.ys = ys
        return self.xhat0 + self.xophat

    def updatestate(self, u, ym, d=None):
        out = SteadyKalmanFilter.updatestate(self, u, ym, d)
        self.xs = self.xs @ self.As_hat.T + self.ys @ self.Bs_hat.T
        return out
