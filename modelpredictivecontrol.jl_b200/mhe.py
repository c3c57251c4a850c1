"""``MovingHorizonEstimator``: host-side mirror of the reference's linear MHE for a BATCH of estimators
(src/estimator/mhe/construct.jl:528-630, ``setconstraint!`` :858-1046, ``preparestate!/updatestate!``
src/estimator/execute.jl:334-386).  LinModel + SingleShooting, ``direct=true`` (the reference default, the
window is solved in ``preparestate``) or ``direct=false`` (prediction form, solved in ``updatestate``).
The per-period work (windows, arrival covariance, Hessian rebuild, QP) runs in libbmpc.so."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, colmajor, dptr
from .host import _b, augment_model


def init_predmat_mhe(A, Bu, Cm, Bd, Ddm, f, He, direct=True):
    """Batched init_predmat_mhe, src/estimator/mhe/transcription.jl:151-260, p = 0 (direct) or 1.
    Block (i, j) formulas (i = 0-based step, j = block column; column 0 = arrival state x̂(k-Nk+p)):
      p = 0:  E[i, arr] = -Cm A^(i+1),  E[i, w_j] = -Cm A^(i-j)  (j <= i),  G[i, u_j] = -Cm A^(i-j) Bu  (j <= i),
              J[i, d_j] = -Cm A^(i-j) Bd (j <= i),  J[i, d_(i+1)] = -Ddm,  B[i] = -Cm S(i) f;
      p = 1:  E[i, arr] = -Cm A^i,      E[i, w_j] = -Cm A^(i-1-j) (j < i),  G[i, u_j] = -Cm A^(i-1-j) Bu (j < i),
              J[i, d_j] = -Cm A^(i-j) Bd (1 <= j <= i),  J[i, d_(i+1)] = -Ddm,
              B[i] = -Cm S(i-1) f for 1 <= i <= He-2 (the reference's own row range, :244-245);
      both:   EX[i, arr] = A^(i+1),  EX[i, w_j] = A^(i-j),  GX[i, u_j] = A^(i-j) Bu,  JX[i, d_(j+p)] = A^(i-j) Bd,
              BX[i] = S(i) f."""
    N, nx = A.shape[0], A.shape[1]
    nu, nym, nd = Bu.shape[2], Cm.shape[1], Bd.shape[2]
    p = 0 if direct else 1
    Ap = [np.broadcast_to(np.eye(nx), (N, nx, nx)).copy()]
    for _ in range(He):
        Ap.append(Ap[-1] @ A)
    S = np.cumsum(np.array(Ap), axis=0)
    nZ = nx * (1 + He)
    E, EX = np.zeros((N, nym * He, nZ)), np.zeros((N, nx * He, nZ))
    G, GX = np.zeros((N, nym * He, nu * He)), np.zeros((N, nx * He, nu * He))
    J, JX = np.zeros((N, nym * He, nd * (He + 1))), np.zeros((N, nx * He, nd * (He + 1)))
    B, BX = np.zeros((N, nym * He)), np.zeros((N, nx * He))
    for i in range(He):
        ry, rx = slice(i * nym, (i + 1) * nym), slice(i * nx, (i + 1) * nx)
        E[:, ry, :nx] = -Cm @ Ap[i + 1 - p]
        EX[:, rx, :nx] = Ap[i + 1]
        if p == 0 or 1 <= i <= He - 2:
            B[:, ry] = -np.einsum("nij,nj->ni", Cm @ S[i - p], f)
        BX[:, rx] = np.einsum("nij,nj->ni", S[i], f)
        if nd:
            J[:, ry, (i + 1) * nd:(i + 2) * nd] = -Ddm
        for j in range(i + 1):
            EX[:, rx, nx + j * nx:nx + (j + 1) * nx] = Ap[i - j]
            GX[:, rx, j * nu:(j + 1) * nu] = Ap[i - j] @ Bu
            if nd:
                JX[:, rx, (j + p) * nd:(j + p + 1) * nd] = Ap[i - j] @ Bd
                if j >= p:
                    J[:, ry, j * nd:(j + 1) * nd] = -Cm @ Ap[i - j] @ Bd
            if j <= i - p:
                E[:, ry, nx + j * nx:nx + (j + 1) * nx] = -Cm @ Ap[i - p - j]
                G[:, ry, j * nu:(j + 1) * nu] = -Cm @ Ap[i - p - j] @ Bu
    return E, G, J, B, EX, GX, JX, BX


class MovingHorizonEstimator:
    def __init__(self, model, He, i_ym=None, sigmaP_0=None, sigmaQ=None, sigmaR=None, nint_u=0, nint_ym=None,
                 sigmaPint_ym_0=None, sigmaQint_ym=None, Cwt=np.inf, shared_model=False, device=0, max_iter=0,
                 tol=0.0, direct=True, P0hat=None, Qhat=None, Rhat=None, transcription="SingleShooting"):
        self.model, self.He, self.direct = model, int(He), bool(direct)
        # transcription = "MultipleShooting" (MovingHorizonEstimator(...; transcription=MultipleShooting()),
        # src/estimator/mhe/construct.jl:546): same estimates -- the equality-constrained problem and the condensed one the
        # kernel solves have the same minimiser --, but `Ztilde` is returned in the reference's MultipleShooting layout
        # [ε; x̂0(k-Nk+p); X̂0; Ŵ] (get_nZ_mhe, mhe/transcription.jl:3), unused entries zero (fill0unused!, :1084-1090)
        if transcription not in ("SingleShooting", "MultipleShooting"):
            raise ValueError("transcription must be 'SingleShooting' or 'MultipleShooting' for a LinModel")
        self.transcription = transcription
        self._nk = 0
        self.__dict__.update(augment_model(model, nint_u, nint_ym, i_ym))
        N, nx, nxh = model.N, model.nx, self.nxhat
        self.nym = len(self.i_ym)
        nsy = self.nxs
        one = lambda v, n, d: np.full(n, d) if v is None else np.asarray(v, float).reshape(n)
        sP = np.concatenate([one(sigmaP_0, nx, 1 / nx), one(sigmaPint_ym_0, nsy, 1.0)])
        sQ = np.concatenate([one(sigmaQ, nx, 1 / nx), one(sigmaQint_ym, nsy, 1.0)])
        sR = one(sigmaR, self.nym, 1.0)
        self.P0hat, self.Qhat, self.Rhat = np.diag(sP ** 2), np.diag(sQ ** 2), np.diag(sR ** 2)
        # full covariance matrices, shared by the batch (the reference's second constructor, mhe/construct.jl:632-660)
        if P0hat is not None: self.P0hat = np.atleast_2d(np.asarray(P0hat, float))
        if Qhat is not None: self.Qhat = np.atleast_2d(np.asarray(Qhat, float))
        if Rhat is not None: self.Rhat = np.atleast_2d(np.asarray(Rhat, float))
        self.Cwt = float(Cwt)
        self.neps = 0 if np.isinf(self.Cwt) else 1
        self.Cmhat, self.Ddmhat = self.Chat[:, self.i_ym], self.Ddhat[:, self.i_ym]
        self.shared_model = bool(shared_model)
        NM = 1 if shared_model else N
        sl = slice(0, NM)
        E, G, J, B, EX, GX, JX, BX = init_predmat_mhe(self.Ahat[sl], self.Buhat[sl], self.Cmhat[sl], self.Bdhat[sl],
                                                      self.Ddmhat[sl], (self.fophat - self.xophat)[sl], self.He,
                                                      self.direct)
        self._h = C.c_void_p()
        dims = _lib.MheDims(N=N, nu=model.nu, nym=self.nym, nd=model.nd, nxhat=nxh, He=self.He, neps=self.neps,
                            direct=int(self.direct), shared_model=int(shared_model), max_iter=max_iter, device=device, tol=tol)
        L = _lib.lib()
        check(L.bmhe_create(C.byref(self._h), C.byref(dims)))
        nd = model.nd
        args = [colmajor(E), colmajor(G), colmajor(J) if nd else None, np.ascontiguousarray(B), colmajor(EX),
                colmajor(GX), colmajor(JX) if nd else None, np.ascontiguousarray(BX)]
        check(L.bmhe_set_predmat(self._h, *[dptr(a) for a in args]))
        rep = lambda M: np.ascontiguousarray(np.broadcast_to(M, (NM,) + M.shape))
        covs = [colmajor(self.Ahat[sl]), colmajor(self.Cmhat[sl]), rep(self.P0hat), rep(self.Qhat), rep(self.Rhat)]
        check(L.bmhe_set_cov(self._h, *[dptr(a) for a in covs], C.c_double(0.0 if np.isinf(self.Cwt) else self.Cwt)))
        inf = np.inf
        self.con = dict(xmin=np.full((N, nxh), -inf), xmax=np.full((N, nxh), inf), wmin=np.full((N, nxh), -inf),
                        wmax=np.full((N, nxh), inf), vmin=np.full((N, self.nym), -inf), vmax=np.full((N, self.nym), inf))
        self.soft = dict(c_x=np.zeros(2 * nxh), c_w=np.zeros(2 * nxh), c_v=np.zeros(2 * self.nym))
        self.xhat0 = np.zeros((N, nxh))
        self._Zss = np.zeros((N, self.neps + nxh * (1 + self.He)))  # the kernel's (SingleShooting) layout [ε; x̂arr; Ŵ]
        self.J = np.zeros(N)
        self.status = np.zeros(N, dtype=np.int32)
        self.iters = np.zeros(N, dtype=np.int32)
        self.Vhat = np.zeros((N, self.nym * self.He))
        self.X0 = np.zeros((N, nxh * self.He))
        self._solved = False

    @property
    def Ztilde(self):
        """The decision vector of the last solve in the reference's layout for the chosen transcription."""
        if self.transcription == "SingleShooting":
            return self._Zss
        N, nxh, He, neps = self.model.N, self.nxhat, self.He, self.neps
        nk = min(self._nk, He)
        Z = np.zeros((N, neps + nxh + 2 * nxh * He))
        nxt = neps + nxh
        Z[:, :nxt] = self._Zss[:, :nxt]
        Z[:, nxt:nxt + nxh * nk] = self.X0[:, :nxh * nk]
        Z[:, nxt + nxh * He:nxt + nxh * He + nxh * nk] = self._Zss[:, nxt:nxt + nxh * nk]
        return Z

    def close(self):
        if self._h:
            _lib.lib().bmhe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setconstraint(self, xhatmin=None, xhatmax=None, whatmin=None, whatmax=None, vhatmin=None, vhatmax=None,
                      c_xhatmin=None, c_xhatmax=None, c_whatmin=None, c_whatmax=None, c_vhatmin=None, c_vhatmax=None):
        N, nxh, nym = self.model.N, self.nxhat, self.nym
        c = self.con
        if xhatmin is not None: c["xmin"] = _b(xhatmin, N, (nxh,), strict=True) - self.xophat
        if xhatmax is not None: c["xmax"] = _b(xhatmax, N, (nxh,), strict=True) - self.xophat
        if whatmin is not None: c["wmin"] = _b(whatmin, N, (nxh,), strict=True).copy()
        if whatmax is not None: c["wmax"] = _b(whatmax, N, (nxh,), strict=True).copy()
        if vhatmin is not None: c["vmin"] = _b(vhatmin, N, (nym,), strict=True).copy()
        if vhatmax is not None: c["vmax"] = _b(vhatmax, N, (nym,), strict=True).copy()
        ecr = [c_xhatmin, c_xhatmax, c_whatmin, c_whatmax, c_vhatmin, c_vhatmax]
        if any(e is not None for e in ecr):
            if not self.neps:
                raise ValueError("Slack variable weight Cwt must be finite to set softness parameters")
            if self._solved:
                raise RuntimeError("Cannot set softness parameters after calling updatestate!")
            s = self.soft

            def w(v, n, name):  # sizes and signs are checked before anything is stored (DimensionMismatch / error)
                v = np.asarray(v, dtype=np.float64).reshape(-1)
                if v.size != n:
                    raise ValueError(f"{name} size must be ({n},)")
                if (v < 0).any():
                    raise ValueError(f"{name} weights should be non-negative")
                return v
            new = [(k, sl, w(v, n, name)) for k, sl, v, n, name in (
                ("c_x", slice(0, nxh), c_xhatmin, nxh, "c_xhatmin"), ("c_x", slice(nxh, 2 * nxh), c_xhatmax, nxh, "c_xhatmax"),
                ("c_w", slice(0, nxh), c_whatmin, nxh, "c_whatmin"), ("c_w", slice(nxh, 2 * nxh), c_whatmax, nxh, "c_whatmax"),
                ("c_v", slice(0, nym), c_vhatmin, nym, "c_vhatmin"), ("c_v", slice(nym, 2 * nym), c_vhatmax, nym, "c_vhatmax"))
                   if v is not None]
            for k, sl, v in new:
                s[k][sl] = v
        a = [np.ascontiguousarray(c[k]) for k in ("xmin", "xmax", "wmin", "wmax", "vmin", "vmax")]
        sv = [np.ascontiguousarray(self.soft[k]) for k in ("c_x", "c_w", "c_v")]
        check(_lib.lib().bmhe_set_constraints(self._h, *[dptr(x) for x in a], *[dptr(x) for x in sv]))
        return self

    def _io(self):
        p32 = lambda a: a.ctypes.data_as(_lib.c_int32_p)
        return [dptr(self.xhat0), dptr(self._Zss), dptr(self.J), p32(self.status), p32(self.iters), dptr(self.Vhat),
                dptr(self.X0)]

    def _dev(self, ym, d):
        m, N = self.model, self.model.N
        y0m = np.ascontiguousarray(_b(ym, N, (self.nym,)) - m.yop[:, self.i_ym])
        d0 = np.ascontiguousarray(_b(d, N, (m.nd,)) - m.dop) if m.nd else None
        return y0m, d0

    def preparestate(self, ym, d=None):
        """preparestate! (src/estimator/execute.jl:334-352): solves the window when direct=true; with direct=false
        correct_estimate! is empty (mhe/execute.jl:44-55) and the current estimate is returned."""
        y0m, d0 = self._dev(ym, d)
        check(_lib.lib().bmhe_correct(self._h, dptr(y0m), dptr(d0) if d0 is not None else None, *self._io()))
        if self.direct:
            self._solved = True
            self._nk += 1
        return self.xhat0 + self.xophat

    def updatestate(self, u, ym=None, d=None):
        """updatestate! (src/estimator/execute.jl:371-386 -> update_estimate!, mhe/execute.jl:71-84)."""
        u0 = np.ascontiguousarray(_b(u, self.model.N, (self.model.nu,)) - self.model.uop)
        if self.direct:
            check(_lib.lib().bmhe_update(self._h, dptr(u0)))
        else:
            if ym is None:
                raise ValueError("updatestate needs ym (and d) when direct=false")
            y0m, d0 = self._dev(ym, d)
            check(_lib.lib().bmhe_update_solve(self._h, dptr(u0), dptr(y0m), dptr(d0) if d0 is not None else None,
                                               *self._io()))
            self._solved = True
            self._nk += 1
        return self.xhat0 + self.xophat

    def reset(self):
        check(_lib.lib().bmhe_reset(self._h))
        self._nk = 0
        self.xhat0[:] = 0
