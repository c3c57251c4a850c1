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


def _controller_arrays(mpcs):
    """Constant data of the oracle controllers in the layout cpuref_linmpc_* reads (instance-major, column-major)."""
    m0 = mpcs[0]
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
    return [E, K, V, B, Ht, Md, con("U0min"), con("U0max"), con("DUmin"), con("DUmax"), con("Y0min"), con("Y0max"),
            sv("C_umin"), sv("C_umax"), sv("C_dumin"), sv("C_dumax"), sv("C_ymin"), sv("C_ymax"),
            st(lambda m: m.model.yop)]


def _dims(m0):
    nu, ny, nx, Hp, Hc, neps = m0.model.nu, m0.model.ny, m0.estim.nxhat, m0.Hp, m0.Hc, m0.neps
    return nu, ny, nx, Hp, Hc, neps, nu * Hc + neps


_p = lambda a: a.ctypes.data_as(C.c_void_p)


def run(mpcs, xhat0, lastu0, ry, threads=0, warm=0):
    """Replay ``steps`` recorded periods on the oracle controllers ``mpcs`` (their matrices and
    constraint state are read; they are not modified).  xhat0/lastu0/ry: (steps, N, len).  The first ``warm`` periods run
    untimed (they warm the solver workspaces).  Returns dict(Z, u, iters, status, seconds, admm_iters, threads)."""
    L = lib()
    N = len(mpcs)
    m0 = mpcs[0]
    nu, ny, nx, Hp, Hc, neps, n = _dims(m0)
    steps = xhat0.shape[0]
    Z = np.zeros((steps, N, n))
    u = np.zeros((steps, N, nu))
    iters = np.zeros((steps, N), dtype=np.int32)
    status = np.zeros((steps, N), dtype=np.int32)
    sec = C.c_double()
    tot = C.c_int64()
    nb = (C.c_int * Hc)(*m0.nb)
    arrs = _controller_arrays(mpcs) + [np.ascontiguousarray(xhat0, dtype=np.float64),
                                       np.ascontiguousarray(lastu0, dtype=np.float64), np.ascontiguousarray(ry, dtype=np.float64)]
    rc = L.cpuref_linmpc_run(C.c_int(N), C.c_int(steps), C.c_int(warm), C.c_int(threads), C.c_int(nu), C.c_int(ny), C.c_int(nx),
                             C.c_int(Hp), C.c_int(Hc), C.c_int(neps), nb, *[_p(a) for a in arrs], _p(Z), _p(u),
                             _p(iters), _p(status), C.byref(sec), C.byref(tot))
    assert rc == 0
    return dict(Z=Z, u=u, iters=iters, status=status, seconds=sec.value, admm_iters=tot.value,
                threads=threads or L.cpuref_num_threads())


def closed_loop(mpcs, ry, threads=0):
    """Closed loop of the oracle controllers ``mpcs`` (plant = their LinModel, their SteadyKalmanFilter as observer,
    operating points zero) entirely on the CPU: records the inputs of every period's moveinput!
    (x̂0 after preparestate!, u0(k-1)) for ``run`` to replay.  ry: (steps, N, ny).  No GPU code is involved."""
    L = lib()
    N = len(mpcs)
    m0 = mpcs[0]
    nu, ny, nx, Hp, Hc, neps, n = _dims(m0)
    nxp = m0.model.nx
    steps = ry.shape[0]
    st = lambda f: np.ascontiguousarray(np.stack([np.asarray(f(m), dtype=np.float64) for m in mpcs]))
    plant = [st(lambda m: m.model.A), st(lambda m: m.model.Bu), st(lambda m: m.model.C)]
    obs = [st(lambda m: m.estim.Ahat), st(lambda m: m.estim.Buhat), st(lambda m: m.estim.Chat), st(lambda m: m.estim.Khat)]
    assert obs[3].shape == (N, nx, ny), "closed_loop expects every output to be measured"
    xh = np.zeros((steps, N, nx))
    lu = np.zeros((steps, N, nu))
    u = np.zeros((steps, N, nu))
    iters = np.zeros((steps, N), dtype=np.int32)
    nb = (C.c_int * Hc)(*m0.nb)
    arrs = _controller_arrays(mpcs) + plant + obs + [np.ascontiguousarray(ry, dtype=np.float64)]
    rc = L.cpuref_linmpc_closed_loop(C.c_int(N), C.c_int(steps), C.c_int(threads), C.c_int(nu), C.c_int(ny), C.c_int(nx),
                                     C.c_int(nxp), C.c_int(Hp), C.c_int(Hc), C.c_int(neps), nb, *[_p(a) for a in arrs],
                                     _p(xh), _p(lu), _p(u), _p(iters))
    assert rc == 0
    return dict(xhat0=xh, lastu0=lu, u=u, iters=iters)


def mhe_run(mhes, windows, threads=0):
    """Timed CPU baseline of the linear MHE on recorded MOVING windows.  ``mhes``: oracle MovingHorizonEstimator
    objects (constant matrices are read from them); ``windows[t][i]`` = the dict ``MovingHorizonEstimator.build_qp()``
    returned for instance i at period t, plus key ``Zs`` (the warm start).  Cwt = Inf (no slack) and a diagonal R̂ are
    assumed, as in BASELINE.json configs[3].  Returns dict(Z, iters, status, seconds, admm_iters, threads)."""
    L = lib()
    N, steps = len(mhes), len(windows)
    w0 = windows[0][0]
    n, nEZ, m = w0["H"].shape[0], w0["EZ"].shape[0], w0["A"].shape[0]
    nxh = mhes[0].nxhat
    assert mhes[0].neps == 0
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    EZ = f(np.stack([windows[0][i]["EZ"] for i in range(N)]))
    Rd = f(np.stack([np.diag(windows[0][i]["Mhat"])[nxh:] for i in range(N)]))
    Nt = f(np.stack([np.diag(windows[0][i]["Nt"]) for i in range(N)]))
    A = f(np.stack([windows[0][i]["A"] for i in range(N)]))
    for t in range(steps):
        for i in range(N):
            assert windows[t][i]["A"].shape == (m, n) and np.array_equal(windows[t][i]["A"], windows[0][i]["A"])
    invP = f(np.stack([[windows[t][i]["Mhat"][:nxh, :nxh] for i in range(N)] for t in range(steps)]))
    FZ = f(np.stack([[windows[t][i]["FZ"] for i in range(N)] for t in range(steps)]))
    b = f(np.stack([[windows[t][i]["b"] for i in range(N)] for t in range(steps)]))
    Zs = f(np.stack([[windows[t][i]["Zs"] for i in range(N)] for t in range(steps)]))
    Z = np.zeros((steps, N, n))
    iters = np.zeros((steps, N), dtype=np.int32)
    status = np.zeros((steps, N), dtype=np.int32)
    sec, tot = C.c_double(), C.c_int64()
    rc = L.cpuref_mhe_run(C.c_int(N), C.c_int(steps), C.c_int(threads), C.c_int(n), C.c_int(nEZ), C.c_int(nxh), C.c_int(m),
                          _p(EZ), _p(Rd), _p(Nt), _p(A), _p(invP), _p(FZ), _p(b), _p(Zs), _p(Z), _p(iters), _p(status),
                          C.byref(sec), C.byref(tot))
    assert rc == 0
    return dict(Z=Z, iters=iters, status=status, seconds=sec.value, admm_iters=tot.value,
                threads=threads or L.cpuref_num_threads())
