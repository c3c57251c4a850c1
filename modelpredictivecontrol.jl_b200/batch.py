"""BatchLinMPC: thin object wrapper over the libbmpc.so handle (one batch of N controllers).

Array convention on the Python side: numpy arrays shaped (N, rows, cols) / (N, len) with the
usual (row, col) meaning; they are converted to the ABI's instance-major column-major layout
here.  With ``shared_model=True`` model-dependent arrays carry a leading dimension of 1.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import BmpcError, check, colmajor, dptr


def _vec(a, N, length, name):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if a.shape == (length,):
        a = np.ascontiguousarray(np.broadcast_to(a, (N, length)))
    if a.shape != (N, length):
        raise ValueError(f"{name} must have shape ({N}, {length}) or ({length},), got {a.shape}")
    return a


class BatchLinMPC:
    """N identically-structured LinMPC controllers stepping in lock-step on one B200."""

    def __init__(self, N, nu, ny, nxhat, Hp, Hc=2, nd=0, Cwt=1e5, shared_model=False, device=0, team=0,
                 max_iter=0, tol=0.0):
        from .host import move_blocking
        self.nb = move_blocking(Hp, Hc)
        self.N, self.nu, self.ny, self.nd, self.nxhat, self.Hp, self.Hc = N, nu, ny, nd, nxhat, Hp, len(self.nb)
        self.neps = 0 if np.isinf(Cwt) else 1
        self.Cwt = float(Cwt)
        self.nDU = nu * self.Hc
        self.n = self.nDU + self.neps
        self.nY, self.nU = ny * Hp, nu * Hp
        self.shared_model = bool(shared_model)
        self.NM = 1 if shared_model else N
        self.device = device
        self._h = C.c_void_p()
        dims = _lib.Dims(N=N, nu=nu, ny=ny, nd=nd, nxhat=nxhat, Hp=Hp, Hc=self.Hc, neps=self.neps,
                         shared_model=int(shared_model), max_iter=max_iter, device=device, team=team, tol=tol)
        nb = (C.c_int32 * self.Hc)(*self.nb)
        check(_lib.lib().bmpc_create(C.byref(self._h), C.byref(dims), nb))
        # controller state owned by the caller, exactly as mpc.Z̃ / mpc.lastu0 in the reference
        self.Ztilde = np.zeros((N, self.n))
        self.lastu0 = np.zeros((N, nu))
        self.u = np.zeros((N, nu))
        self.J = np.zeros(N)
        self.status = np.zeros(N, dtype=np.int32)
        self.iters = np.zeros(N, dtype=np.int32)
        self.kkt = np.zeros((N, 3))  # relative KKT residuals of the returned iterate (primal, dual, complementarity)
        self.nw = 0
        self.uop = np.zeros((self.NM, nu))
        self.yop = np.zeros((self.NM, ny))

    def close(self):
        if self._h:
            _lib.lib().bmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- model-dependent constants ---------------------------------------------------------
    def _mat(self, a, rows, cols, name):
        if a is None:
            return None
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 2:
            a = a[None]
        if a.shape != (self.NM, rows, cols):
            raise ValueError(f"{name} must have shape ({self.NM}, {rows}, {cols}), got {a.shape}")
        return colmajor(a)

    def set_predmat(self, E, K, V, B, Htilde, G=None, J=None, ex=None, kx=None, vx=None, bx=None, gx=None, jx=None):
        """Route B: host-computed mpc.Ẽ (without slack column), K, V, B, G, J, H̃ and terminal matrices."""
        nY, nDU, nx, nu, nd, Hp, n = self.nY, self.nDU, self.nxhat, self.nu, self.nd, self.Hp, self.n
        v = lambda a, ln, nm: None if a is None else _vec(a, self.NM, ln, nm)
        args = [self._mat(E, nY, nDU, "E"), self._mat(K, nY, nx, "K"), self._mat(V, nY, nu, "V"), v(B, nY, "B"),
                self._mat(G, nY, nd, "G") if nd else None, self._mat(J, nY, nd * Hp, "J") if nd else None,
                self._mat(Htilde, n, n, "Htilde"), self._mat(ex, nx, nDU, "ex"), self._mat(kx, nx, nx, "kx"),
                self._mat(vx, nx, nu, "vx"), v(bx, nx, "bx"), self._mat(gx, nx, nd, "gx") if nd else None,
                self._mat(jx, nx, nd * Hp, "jx") if nd else None]
        check(_lib.lib().bmpc_set_predmat(self._h, *[dptr(a) for a in args]))

    def set_model(self, Ahat, Buhat, Chat, Bdhat=None, Ddhat=None, fop_minus_xop=None, M_diag=None, N_diag=None,
                  L_diag=None):
        """Route A: augmented model + diagonal weights; prediction matrices and Hessian are built on the GPU."""
        nx, nu, ny, nd = self.nxhat, self.nu, self.ny, self.nd
        NM = self.NM
        f = np.zeros((NM, nx)) if fop_minus_xop is None else _vec(fop_minus_xop, NM, nx, "fop_minus_xop")
        M = _vec(np.ones(self.nY) if M_diag is None else M_diag, NM, self.nY, "M_diag")
        Nw = _vec(np.full(self.nDU, 0.1) if N_diag is None else N_diag, NM, self.nDU, "N_diag")
        L = _vec(np.zeros(self.nU) if L_diag is None else L_diag, NM, self.nU, "L_diag")
        args = [self._mat(Ahat, nx, nx, "Ahat"), self._mat(Buhat, nx, nu, "Buhat"), self._mat(Chat, ny, nx, "Chat"),
                self._mat(Bdhat, nx, nd, "Bdhat") if nd else None, self._mat(Ddhat, ny, nd, "Ddhat") if nd else None,
                f, M, Nw, L]
        check(_lib.lib().bmpc_set_model(self._h, *[dptr(a) for a in args], C.c_double(self.Cwt)))

    def set_weights(self, M, L_diag=None):
        M = np.asarray(M, dtype=np.float64)
        dense = M.ndim >= 2 and M.shape[-1] == self.nY and M.shape[-2] == self.nY and self.nY > 1
        if dense:
            Mc = self._mat(M, self.nY, self.nY, "M_Hp")
        else:
            Mc = _vec(M, self.NM, self.nY, "M_diag")
        Ld = False
        if L_diag is None:
            L = None
        else:
            La = np.asarray(L_diag, dtype=np.float64)
            Ld = La.ndim >= 2 and La.shape[-1] == self.nU and La.shape[-2] == self.nU and self.nU > 1
            L = self._mat(La, self.nU, self.nU, "L_Hp") if Ld else _vec(La, self.NM, self.nU, "L_diag")
        check(_lib.lib().bmpc_set_weights_dense(self._h, dptr(Mc), int(dense), dptr(L), int(Ld)))

    def set_oppoints(self, uop=None, yop=None):
        if uop is not None:
            self.uop = _vec(uop, self.NM, self.nu, "uop")
        if yop is not None:
            self.yop = _vec(yop, self.NM, self.ny, "yop")
        check(_lib.lib().bmpc_set_oppoints(self._h, dptr(self.uop), dptr(self.yop)))

    def set_constraints(self, U0min=None, U0max=None, DUmin=None, DUmax=None, Y0min=None, Y0max=None,
                        xhat0min=None, xhat0max=None, soft=None):
        """Bounds in deviation form (the content of mpc.con after setconstraint!); None = unbounded."""
        N = self.N
        v = lambda a, ln, nm: None if a is None else _vec(a, N, ln, nm)
        arrs = [v(U0min, self.nU, "U0min"), v(U0max, self.nU, "U0max"), v(DUmin, self.nDU, "DUmin"),
                v(DUmax, self.nDU, "DUmax"), v(Y0min, self.nY, "Y0min"), v(Y0max, self.nY, "Y0max"),
                v(xhat0min, self.nxhat, "xhat0min"), v(xhat0max, self.nxhat, "xhat0max")]
        S = _lib.Softness()
        keep = []
        if soft:
            lens = dict(C_umin=self.nU, C_umax=self.nU, C_dumin=self.nDU, C_dumax=self.nDU, C_ymin=self.nY,
                        C_ymax=self.nY, c_xmin=self.nxhat, c_xmax=self.nxhat)
            for k, val in soft.items():
                if val is None:
                    continue
                a = np.ascontiguousarray(np.asarray(val, dtype=np.float64).reshape(lens[k]))
                keep.append(a)
                setattr(S, k, dptr(a))
        check(_lib.lib().bmpc_set_constraints(self._h, *[dptr(a) for a in arrs], C.byref(S)))

    def set_custom(self, nw, Wy=None, Wu=None, Wd=None, Wr=None, Chat=None, Ddhat=None, dop=None):
        """Custom linear constraints (bmpc_set_custom): matrices (NM, nw, cols) or (nw, cols); Chat / Ddhat of the estimator;
        dop per model.  Call before ``set_constraints``."""
        nx, nu, ny, nd, NM = self.nxhat, self.nu, self.ny, self.nd, self.NM
        self.nw = int(nw)
        if nw == 0:
            check(_lib.lib().bmpc_set_custom(self._h, 0, *([None] * 7)))
            return
        m = lambda a, r, c, nm: None if a is None else self._mat(np.broadcast_to(np.asarray(a, float).reshape(-1, r, c), (NM, r, c)), r, c, nm)
        args = [m(Wy, nw, ny, "Wy"), m(Wu, nw, nu, "Wu"), m(Wd, nw, nd, "Wd") if nd else None, m(Wr, nw, ny, "Wr"),
                self._mat(Chat, ny, nx, "Chat"), self._mat(Ddhat, ny, nd, "Ddhat") if nd else None,
                _vec(dop, NM, nd, "dop") if (nd and dop is not None) else None]
        check(_lib.lib().bmpc_set_custom(self._h, int(nw), *[dptr(a) for a in args]))

    def set_custom_bounds(self, Wmin=None, Wmax=None, C_wmin=None, C_wmax=None):
        """Bounds of the custom rows, absolute units, (N, nw (Hp + 1)); compiled by the next ``set_constraints``."""
        nFw = self.nw * (self.Hp + 1)
        v = lambda a: None if a is None else _vec(a, self.N, nFw, "W bound")
        c = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, float).reshape(nFw))
        args = [v(Wmin), v(Wmax), c(C_wmin), c(C_wmax)]
        check(_lib.lib().bmpc_set_custom_bounds(self._h, *[dptr(a) for a in args]))

    def set_estimator(self, Ahat, Buhat, Cmhat, Khat, Bdhat=None, Ddmhat=None, fop_minus_xop=None):
        """Fused SteadyKalmanFilter (bmpc_set_estimator): the handle owns x̂0; ``step(None, y0m=...)`` then runs
        correct -> moveinput! -> predict in one launch."""
        nx, nu, nd = self.nxhat, self.nu, self.nd
        nym = np.asarray(Cmhat).shape[-2]
        fx = None if fop_minus_xop is None else _vec(fop_minus_xop, self.NM, nx, "fop_minus_xop")
        args = [self._mat(Ahat, nx, nx, "Ahat"), self._mat(Buhat, nx, nu, "Buhat"),
                self._mat(Bdhat, nx, nd, "Bdhat") if nd else None, self._mat(Cmhat, nym, nx, "Cmhat"),
                self._mat(Ddmhat, nym, nd, "Ddmhat") if nd else None,
                None if Khat is None else self._mat(Khat, nx, nym, "Khat"), fx]
        check(_lib.lib().bmpc_set_estimator(self._h, *[dptr(a) for a in args], int(nym)))
        self.nym = nym

    def set_estimator_cov(self, P0, Qhat, Rhat):
        """Time-varying KalmanFilter as the fused observer (bmpc_set_estimator_cov): P̂_0, Q̂, R̂ per model; the gain is
        recomputed every period by the covariance recursion around the step kernel."""
        nx, nym = self.nxhat, self.nym
        self._kf = [self._mat(P0, nx, nx, "P0"), self._mat(Qhat, nx, nx, "Qhat"), self._mat(Rhat, nym, nym, "Rhat")]
        check(_lib.lib().bmpc_set_estimator_cov(self._h, *[dptr(a) for a in self._kf]))

    def get_cov(self):
        P = np.zeros((self.NM, self.nxhat, self.nxhat))
        check(_lib.lib().bmpc_get_cov(self._h, dptr(P)))
        return np.swapaxes(P, 1, 2)  # column-major on the device

    def set_state(self, xhat0):
        check(_lib.lib().bmpc_set_state(self._h, dptr(_vec(xhat0, self.N, self.nxhat, "xhat0"))))

    def get_state(self):
        """(x̂0 predicted for the next period, corrected x̂0 the last step used)."""
        a, b = np.zeros((self.N, self.nxhat)), np.zeros((self.N, self.nxhat))
        check(_lib.lib().bmpc_get_state(self._h, dptr(a), dptr(b)))
        return a, b

    # ---- per-period call (= moveinput!) ----------------------------------------------------
    def step(self, xhat0, ry=None, Rhat_y=None, Rhat_u=None, d0=None, Dhat0=None, resident=False, y0m=None, Yhat_s=None):
        """One control period (moveinput!).  ``resident=True``: lastu0 and Z̃ stay in the handle between calls
        (they are fields of the reference controller, not arguments of moveinput!): only x̂0/ry go up and only
        u/status come back; self.lastu0, self.Ztilde, self.J, self.iters are then NOT refreshed."""
        N = self.N
        x = None if xhat0 is None else _vec(xhat0, N, self.nxhat, "xhat0")
        ym = None if y0m is None else _vec(y0m, N, self.nym, "y0m")
        ryv = None if ry is None else _vec(ry, N, self.ny, "ry")
        Ry = None if Rhat_y is None else _vec(Rhat_y, N, self.nY, "Rhat_y")
        Ru = None if Rhat_u is None else _vec(Rhat_u, N, self.nU, "Rhat_u")
        d0v = None if d0 is None or self.nd == 0 else _vec(d0, N, self.nd, "d0")
        Dh = None if Dhat0 is None or self.nd == 0 else _vec(Dhat0, N, self.nd * self.Hp, "Dhat0")
        Ys = None if Yhat_s is None else _vec(Yhat_s, N, self.nY, "Yhat_s")
        p = lambda a: None if a is None else a.ctypes.data
        if resident:
            io = _lib.StepIO(xhat0=p(x), ry=p(ryv), Rhat_y=p(Ry), Rhat_u=p(Ru), d0=p(d0v), Dhat0=p(Dh), u=p(self.u),
                             status=p(self.status), device_ptrs=0, sync=1, resident=1, y0m=p(ym), Yhat_s=p(Ys), kkt=p(self.kkt))
        else:
            io = _lib.StepIO(xhat0=p(x), lastu0=p(self.lastu0), ry=p(ryv), Rhat_y=p(Ry), Rhat_u=p(Ru), d0=p(d0v),
                             Dhat0=p(Dh), Ztilde=p(self.Ztilde), u=p(self.u), J=p(self.J), status=p(self.status),
                             iters=p(self.iters), device_ptrs=0, sync=1, y0m=p(ym), Yhat_s=p(Ys), kkt=p(self.kkt))
        check(_lib.lib().bmpc_step(self._h, C.byref(io)))
        return self.u

    def step_mapped(self, xhat0, ry, u, status, resident=True, y0m=None):
        """Zero-copy call (io.host_mapped = 1): every array must be PAGE-LOCKED host memory that the device can address
        (e.g. ``torch.Tensor.pin_memory().numpy()``); the kernel reads x̂0/ry from and writes u/status to it directly."""
        p = lambda a: None if a is None else a.ctypes.data
        kw = dict(xhat0=p(xhat0)) if y0m is None else dict(y0m=p(y0m))
        if resident:
            io = _lib.StepIO(ry=p(ry), u=p(u), status=p(status), device_ptrs=0, sync=1, resident=1, host_mapped=1, **kw)
        else:
            # lastu0 / Z̃ go through staging copies (pageable memory is fine); iters would have to be pinned: omitted
            io = _lib.StepIO(lastu0=p(self.lastu0), ry=p(ry), Ztilde=p(self.Ztilde), u=p(u), status=p(status),
                             device_ptrs=0, sync=1, host_mapped=1, **kw)
        check(_lib.lib().bmpc_step(self._h, C.byref(io)))
        return u

    def step_device(self, ptrs, sync=False):
        """Device-pointer variant: ``ptrs`` maps StepIO field names to raw device addresses (ints)."""
        io = _lib.StepIO(device_ptrs=1, sync=int(sync), **ptrs)
        check(_lib.lib().bmpc_step(self._h, C.byref(io)))

    def set_gather(self, peer_ptrs, rank):
        """Fused all-gather of Z̃: ``peer_ptrs[p]`` = device address (int) of rank p's [world, N, n] float64 gather
        buffer mapped into this process (symmetric memory); the step kernel then stores Z̃ into every peer's buffer."""
        world = len(peer_ptrs) if peer_ptrs else 0
        arr = (C.c_void_p * max(world, 1))(*[C.c_void_p(int(p)) for p in (peer_ptrs or [])])
        check(_lib.lib().bmpc_set_gather(self._h, arr if world else None, world, int(rank)))

    def set_gather_pull(self, peer_ptrs, flag_ptrs, rank, row_offsets, slots=3):
        """Fused all-gather of Z̃, pull protocol (include/bmpc.h bmpc_set_gather_pull): ``peer_ptrs[p]`` / ``flag_ptrs[p]`` =
        device addresses of rank p's [slots, N_p, n] float64 slot buffer and [2 world] uint64 flag array, mapped into this
        process; ``row_offsets`` (world + 1) places every rank's rows in the gathered array."""
        world = len(peer_ptrs)
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        farr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in flag_ptrs])
        ro = (C.c_int32 * (world + 1))(*[int(v) for v in row_offsets])
        check(_lib.lib().bmpc_set_gather_pull(self._h, arr, farr, world, int(rank), ro, int(slots)))

    def gather_epoch(self):
        return int(_lib.lib().bmpc_gather_epoch(self._h))

    def gather_pull(self, epoch, dst_ptr, stream_ptr=None):
        """Enqueue the one-sided gather of period ``epoch`` into the device array at ``dst_ptr`` ([rows_total, n])."""
        check(_lib.lib().bmpc_gather_pull(self._h, int(epoch), C.c_void_p(int(dst_ptr)) if dst_ptr else None,
                                          C.c_void_p(int(stream_ptr)) if stream_ptr else None))

    def gather_timed_out(self):
        return int(_lib.lib().bmpc_gather_timed_out(self._h))

    def get_states(self):
        """X̂0 block of the MultipleShooting decision vector for the last step (bmpc_get_states): (N, nxhat Hp)."""
        X0 = np.zeros((self.N, self.nxhat * self.Hp))
        check(_lib.lib().bmpc_get_states(self._h, dptr(X0)))
        return X0

    def set_stream(self, stream_ptr):
        check(_lib.lib().bmpc_set_stream(self._h, C.c_void_p(stream_ptr)))

    def getinfo(self):
        N = self.N
        out = dict(Yhat0=np.zeros((N, self.nY)), U0=np.zeros((N, self.nU)), xhat0end=np.zeros((N, self.nxhat)),
                   F=np.zeros((N, self.nY)), qtilde=np.zeros((N, self.n)), r=np.zeros(N))
        info = _lib.Info(**{k: v.ctypes.data for k, v in out.items()})
        check(_lib.lib().bmpc_getinfo(self._h, C.byref(info)))
        out.update(DU=self.Ztilde[:, :self.nDU].copy(), eps=self.Ztilde[:, -1].copy() if self.neps else np.zeros(N),
                   J=self.J.copy(), status=self.status.copy(), iters=self.iters.copy())
        return out

    def launch_info(self):
        out = (C.c_int32 * 8)()
        check(_lib.lib().bmpc_launch_info(self._h, out))
        keys = ("team", "teams_per_cta", "grid", "smem_bytes_per_cta", "pd_in_smem", "rows_m", "sparse_rows", "dense_rows")
        return dict(zip(keys, list(out)))

    def launch_count(self):
        return int(_lib.lib().bmpc_launch_count(self._h))
