"""ctypes wrapper of the CPU baseline (restated reference path: assembly + OSQP-style ADMM).
TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py and linmpc_admm.cpp)."""
import ctypes as C
import os

import numpy as np

from .build import OUT, build

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(OUT):
            build()
        _lib = C.CDLL(OUT)
        _lib.cpuref_num_threads.restype = C.c_int
    return _lib


def run(mpcs, xhat0, lastu0, ry, threads=0):
    """Replay ``steps`` recorded periods on the oracle controllers ``mpcs`` (their matrices and
    constraint state are read; they are not modified).  xhat0/lastu0/ry: (steps, N, len).
    Returns dict(Z, u, iters, status, seconds, admm_iters, threads)."""
    L = lib()
    N = len(mpcs)
    m0 = mpcs[0]
    nu, ny, nx, Hp, Hc, neps = m0.model.nu, m0.model.ny, m0.estim.nxhat, m0.Hp, m0.Hc, m0.neps
    nz = nu * Hc
    n = nz + neps
    steps = xhat0.shape[0]
    cm = lambda a: np.ascontiguousarray(np.swapaxes(np.asarray(a, dtype=np.float64), -1, -2))
    st = lambda f: np.ascontiguousarray(np.stack([f(m) for m in mpcs]), dtype=np.float64)
    E, K, V = cm(st(lambda m: m.E)), cm(st(lambda m: m.K)), cm(st(lambda m: m.V))
    B, Ht = st(lambda m: m.B), cm(st(lambda m: m.Htilde))
    Md = st(lambda m: np.diag(m.M_Hp))
    con = lambda k: st(lambda m: getattr(m.con, k))
    c0 = m0.con
    sv = lambda k: np.ascontiguousarray(getattr(c0, k), dtype=np.float64)
    if any(np.isfinite(m.con.xhat0min).any() or np.isfinite(m.con.xhat0max).any() for m in mpcs):
        raise NotImplementedError("terminal constraints are not part of the CPU baseline")
    Z = np.zeros((steps, N, n))
    u = np.zeros((steps, N, nu))
    iters = np.zeros((steps, N), dtype=np.int32)
    status = np.zeros((steps, N), dtype=np.int32)
    sec = C.c_double()
    tot = C.c_int64()
    nb = (C.c_int * Hc)(*m0.nb)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    arrs = [E, K, V, B, Ht, Md, con("U0min"), con("U0max"), con("DUmin"), con("DUmax"), con("Y0min"), con("Y0max"),
            sv("C_umin"), sv("C_umax"), sv("C_dumin"), sv("C_dumax"), sv("C_ymin"), sv("C_ymax"),
            st(lambda m: m.model.yop), np.ascontiguousarray(xhat0, dtype=np.float64),
            np.ascontiguousarray(lastu0, dtype=np.float64), np.ascontiguousarray(ry, dtype=np.float64)]
    rc = L.cpuref_linmpc_run(C.c_int(N), C.c_int(steps), C.c_int(threads), C.c_int(nu), C.c_int(ny), C.c_int(nx),
                             C.c_int(Hp), C.c_int(Hc), C.c_int(neps), nb, *[p(a) for a in arrs], p(Z), p(u),
                             p(iters), p(status), C.byref(sec), C.byref(tot))
    assert rc == 0
    return dict(Z=Z, u=u, iters=iters, status=status, seconds=sec.value, admm_iters=tot.value,
                threads=threads or L.cpuref_num_threads())
