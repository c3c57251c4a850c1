"""GPU parity tests of the batched linear MovingHorizonEstimator (bmhe_* through the C ABI) against
oracle/mhe.py: growing and moving windows, arrival-covariance recursion, per-step Hessian rebuild,
hard and soft bounds, measured disturbance + operating points.  Tolerances: states and Z̃ 1e-6
relative when constraints are active (IPM), 1e-9 when not (Cholesky exit); J 1e-8 relative."""
import numpy as np
import pytest

from oracle.linmpc import LinModel as OLinModel
from oracle.mhe import KalmanFilter as OKF, MovingHorizonEstimator as OMHE

pytestmark = pytest.mark.gpu


def make(N, seed, nx=3, nu=2, ny=2, nd=1):
    import mpc_b200
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((N, nx, nx))
    A *= (rng.uniform(0.5, 0.9, N) / np.abs(np.linalg.eigvals(A)).max(axis=1))[:, None, None]
    Bu, Cc = rng.standard_normal((N, nx, nu)), rng.standard_normal((N, ny, nx))
    Bd, Dd = rng.standard_normal((N, nx, nd)), 0.1 * rng.standard_normal((N, ny, nd))
    op = dict(uop=[1.0, -2.0][:nu], yop=[5.0, 3.0][:ny], dop=[0.5][:nd])
    gm = mpc_b200.LinModel(A, Bu, Cc, Bd=Bd if nd else None, Dd=Dd if nd else None, N=N, **op)
    oms = [OLinModel(A[i], Bu[i], Cc[i], Bd=Bd[i] if nd else None, Dd=Dd[i] if nd else None, **op) for i in range(N)]
    return gm, oms, rng


def run_both(gmhe, omhes, rng, steps, nd, ny, nu, tol_active=1e-6):
    N = len(omhes)
    worst = 0.0
    nact = 0
    if not gmhe.direct:
        return run_both_prediction_form(gmhe, omhes, rng, steps, nd, ny, nu, tol_active)
    for k in range(steps):
        y = np.array([5.0, 3.0])[:ny] + rng.standard_normal((N, ny))
        d = 0.5 + 0.3 * rng.standard_normal((N, nd))
        u = np.array([1.0, -2.0])[:nu] + rng.standard_normal((N, nu))
        xg = gmhe.preparestate(y, d if nd else None)
        for i, o in enumerate(omhes):
            xo = o.preparestate(y[i], d[i] if nd else ())
            assert gmhe.status[i] == o.last_qp["status"], (k, i, gmhe.status[i], o.last_qp["status"], gmhe.iters[i])
            tol = tol_active if gmhe.iters[i] > 0 else 1e-9
            nact += gmhe.iters[i] > 0
            e = np.abs(xg[i] - xo).max() / (1 + np.abs(xo).max())
            ez = np.abs(gmhe.Ztilde[i] - o.Ztilde).max() / (1 + np.abs(o.Ztilde).max())
            ej = abs(gmhe.J[i] - o.Jval) / (1 + abs(o.Jval))
            assert e < tol and ez < tol and ej < 1e-8, (k, i, e, ez, ej, gmhe.iters[i])
            worst = max(worst, e, ez)
            o.updatestate(u[i], y[i], d[i] if nd else ())
        gmhe.updatestate(u, y, d if nd else None)
    return worst, nact


def run_both_prediction_form(gmhe, omhes, rng, steps, nd, ny, nu, tol_active):
    """direct=false: preparestate! returns the current estimate unchanged, the window is solved in updatestate!
    (update_estimate!, src/estimator/mhe/execute.jl:71-84)."""
    N = len(omhes)
    worst, nact = 0.0, 0
    for k in range(steps):
        y = np.array([5.0, 3.0])[:ny] + rng.standard_normal((N, ny))
        d = 0.5 + 0.3 * rng.standard_normal((N, nd))
        u = np.array([1.0, -2.0])[:nu] + rng.standard_normal((N, nu))
        xp = gmhe.preparestate(y, d if nd else None).copy()
        for i, o in enumerate(omhes):
            xo = o.preparestate(y[i], d[i] if nd else ())
            assert np.abs(xp[i] - xo).max() <= 2e-6 * (1 + np.abs(xo).max()), (k, i)
        xg = gmhe.updatestate(u, y, d if nd else None)
        for i, o in enumerate(omhes):
            xo = o.updatestate(u[i], y[i], d[i] if nd else ())
            assert gmhe.status[i] == o.last_qp["status"], (k, i, gmhe.status[i], o.last_qp["status"], gmhe.iters[i])
            tol = tol_active if gmhe.iters[i] > 0 else 1e-9
            nact += gmhe.iters[i] > 0
            e = np.abs(xg[i] - xo).max() / (1 + np.abs(xo).max())
            ez = np.abs(gmhe.Ztilde[i] - o.Ztilde).max() / (1 + np.abs(o.Ztilde).max())
            ej = abs(gmhe.J[i] - o.Jval) / (1 + abs(o.Jval))
            assert e < tol and ez < tol and ej < 1e-8, (k, i, e, ez, ej, gmhe.iters[i])
            worst = max(worst, e, ez)
    return worst, nact


@pytest.mark.parametrize("He,nd", [(3, 1), (5, 0), (1, 1)])
def test_mhe_prediction_form_matches_oracle_and_kalman(He, nd):
    """test/2_test_state_estim.jl:1750-1766 through the GPU: MHE (direct=false) == oracle MHE == KalmanFilter
    (direct=false); covers the growing window, the first full window and the moving window with the
    correct+predict covariance update."""
    import mpc_b200
    N = 6
    gm, oms, rng = make(N, 4, nd=nd)
    g = mpc_b200.MovingHorizonEstimator(gm, He=He, nint_ym=[0, 0], direct=False)
    os_ = [OMHE(m, He=He, nint_ym=0, direct=False) for m in oms]
    kfs = [OKF(m, nint_ym=0, direct=False) for m in oms]
    worst, nact = run_both(g, os_, rng, 2 * He + 4, nd, 2, 2)
    assert nact == 0
    # the same data through the time-varying KalmanFilter (oracle side already asserted equal in
    # tests/test_oracle_mhe.py); here: GPU MHE state == KF state after the same sequence
    rng2 = np.random.default_rng(4)
    gm2, oms2, rng2 = make(N, 4, nd=nd)
    for k in range(2 * He + 4):
        y = np.array([5.0, 3.0]) + rng2.standard_normal((N, 2))
        d = 0.5 + 0.3 * rng2.standard_normal((N, nd))
        u = np.array([1.0, -2.0]) + rng2.standard_normal((N, 2))
        for i, kf in enumerate(kfs):
            kf.preparestate(y[i], d[i] if nd else ())
            kf.updatestate(u[i], y[i], d[i] if nd else ())
    xk = np.stack([kf.xhat0 + kf.xophat for kf in kfs])
    xg = g.xhat0 + g.xophat
    assert np.abs(xg - xk).max() < 1e-6 * (1 + np.abs(xk).max()), np.abs(xg - xk).max()
    print("MHE direct=false worst", worst)


def test_mhe_prediction_form_bounds_match_oracle():
    """direct=false with state / process-noise / sensor-noise bounds, hard."""
    import mpc_b200
    N, He = 5, 4
    gm, oms, rng = make(N, 6, nd=1)
    kw = dict(xhatmin=[-0.6] * 3, xhatmax=[0.6] * 3, whatmin=[-0.3] * 3, whatmax=[0.3] * 3,
              vhatmin=[-2.5] * 2, vhatmax=[2.5] * 2)
    g = mpc_b200.MovingHorizonEstimator(gm, He=He, nint_ym=[0, 0], direct=False).setconstraint(**kw)
    os_ = [OMHE(m, He=He, nint_ym=0, direct=False).setconstraint(**kw) for m in oms]
    worst, nact = run_both(g, os_, rng, 2 * He + 3, 1, 2, 2, tol_active=2e-6)
    assert nact > 10
    print("MHE direct=false constrained worst", worst, "active solves", nact)


@pytest.mark.parametrize("He,nd", [(3, 1), (5, 0)])
def test_mhe_unconstrained_matches_oracle_and_kalman(He, nd):
    """test/2_test_state_estim.jl:1767-1784 through the GPU: MHE (direct=true) == oracle MHE == KalmanFilter."""
    import mpc_b200
    N = 6
    gm, oms, rng = make(N, 3, nd=nd)
    g = mpc_b200.MovingHorizonEstimator(gm, He=He, nint_ym=[0, 0])
    os_ = [OMHE(m, He=He, nint_ym=0) for m in oms]
    worst, nact = run_both(g, os_, rng, 2 * He + 4, nd, 2, 2)
    assert nact == 0
    print("MHE unconstrained worst", worst)


@pytest.mark.parametrize("Cwt", [np.inf, 1e5])
def test_mhe_bounds_match_oracle(Cwt):
    """State / process-noise / sensor-noise bounds (config C3 recipe at small size): hard and soft."""
    import mpc_b200
    N, He = 5, 4
    gm, oms, rng = make(N, 5, nd=1)
    kw = dict(xhatmin=[-0.6] * 3, xhatmax=[0.6] * 3, whatmin=[-0.3] * 3, whatmax=[0.3] * 3,
              vhatmin=[-2.5] * 2, vhatmax=[2.5] * 2)
    soft = dict(c_xhatmin=[1] * 3, c_xhatmax=[1] * 3, c_whatmin=[0.1] * 3, c_whatmax=[0.1] * 3,
                c_vhatmin=[1] * 2, c_vhatmax=[1] * 2) if np.isfinite(Cwt) else {}
    g = mpc_b200.MovingHorizonEstimator(gm, He=He, nint_ym=[0, 0], Cwt=Cwt).setconstraint(**kw, **soft)
    os_ = [OMHE(m, He=He, nint_ym=0, Cwt=Cwt).setconstraint(**kw, **soft) for m in oms]
    worst, nact = run_both(g, os_, rng, 2 * He + 3, 1, 2, 2, tol_active=2e-6)
    assert nact > 10
    print("MHE constrained worst", worst, "active solves", nact)


def test_mhe_with_integrators_shared_model():
    """nint_ym = 1 per output (augmented states) and one model shared by all instances."""
    import mpc_b200
    N, He = 4, 3
    gm1, oms, rng = make(1, 9, nd=0)
    gm = mpc_b200.LinModel(gm1.A[0], gm1.Bu[0], gm1.C[0], N=N, uop=[1.0, -2.0], yop=[5.0, 3.0])
    g = mpc_b200.MovingHorizonEstimator(gm, He=He, shared_model=True)
    os_ = [OMHE(oms[0], He=He) for _ in range(N)]
    worst, _ = run_both(g, os_, rng, 2 * He + 2, 0, 2, 2)
    print("MHE integrators worst", worst)


def test_mhe_status_agreement_many_windows():
    """Feasible / infeasible classification on many windows: hard bounds tight enough that a large share of the windows
    is infeasible (sensor noise against |v̂| <= 2); the GPU status (two-collapsed-steps exit, Farkas-ray checkpoints)
    must equal the exact oracle's on every window, and feasible windows must agree on x̂ (5e-6 / 1e-9)."""
    import mpc_b200
    N, He = 40, 6
    gm, oms, rng = make(N, 31, nx=4, nd=0)
    kw = dict(xhatmin=[-10] * 4, xhatmax=[10] * 4, whatmin=[-0.5] * 4, whatmax=[0.5] * 4, vhatmin=[-2.0] * 2,
              vhatmax=[2.0] * 2)
    g = mpc_b200.MovingHorizonEstimator(gm, He=He, nint_ym=[0, 0]).setconstraint(**kw)
    os_ = [OMHE(m, He=He, nint_ym=0).setconstraint(**kw) for m in oms]
    x = np.zeros((N, 4))
    ninf = nfeas_active = 0
    for k in range(14):
        u = rng.choice([-1.0, 1.0], (N, 2))
        x = np.einsum("nij,nj->ni", gm.A, x) + np.einsum("nij,nj->ni", gm.Bu, u - gm.uop) + rng.standard_normal((N, 4)) / 4
        y = np.einsum("nij,nj->ni", gm.C, x) + gm.yop + rng.standard_normal((N, 2))
        xg = g.preparestate(y).copy()
        for i, o in enumerate(os_):
            xo = o.preparestate(y[i])
            assert g.status[i] == o.last_qp["status"], (k, i, g.status[i], o.last_qp["status"], g.iters[i])
            if g.status[i] == 0:
                tol = 5e-6 if g.iters[i] > 0 else 1e-9
                assert np.abs(xg[i] - xo).max() < tol * (1 + np.abs(xo).max()), (k, i, g.iters[i])
                nfeas_active += int(g.iters[i] > 0)
            else:
                ninf += 1
            o.updatestate(u[i], y[i])
        g.updatestate(u, y)
    assert ninf > 40 and nfeas_active > 40, (ninf, nfeas_active)
    print("MHE status agreement: infeasible windows", ninf, "feasible active windows", nfeas_active, "of", N * 14)


def test_mhe_nan_measurements_match_oracle():
    """Missing measurements (NaN in ym): the reference zeroes the affected rows of Ẽ and of F before building H̃ / q̃
    (src/estimator/mhe/execute.jl:436-441); the kernel does the same (bmpc_mhe.cuh).  Some instances lose one output for
    one or two periods inside the window; estimates must follow the oracle through the growing and moving windows."""
    import mpc_b200
    N, He = 6, 4
    gm, oms, rng = make(N, 13, nd=0)
    g = mpc_b200.MovingHorizonEstimator(gm, He=He, nint_ym=[0, 0])
    os_ = [OMHE(m, He=He, nint_ym=0) for m in oms]
    worst = 0.0
    for k in range(2 * He + 4):
        y = np.array([5.0, 3.0]) + rng.standard_normal((N, 2))
        u = np.array([1.0, -2.0]) + rng.standard_normal((N, 2))
        if k in (2, 3, 7):
            y[k % N, k % 2] = np.nan       # one output of one instance is missing this period
        if k == 6:
            y[1, :] = np.nan               # both outputs of instance 1
        xg = g.preparestate(y)
        for i, o in enumerate(os_):
            xo = o.preparestate(y[i])
            assert np.isfinite(xg[i]).all(), (k, i)
            e = np.abs(xg[i] - xo).max() / (1 + np.abs(xo).max())
            ej = abs(g.J[i] - o.Jval) / (1 + abs(o.Jval))
            assert e < 1e-9 and ej < 1e-8, (k, i, e, ej)
            worst = max(worst, e)
            o.updatestate(u[i], y[i])
        g.updatestate(u, y)
    print("MHE NaN measurements worst", worst)


def test_mhe_dense_covariances_match_oracle():
    """Non-diagonal R̂ (and Q̂, P̂_0): the reference's full-matrix constructor (src/estimator/mhe/construct.jl:632-660).
    invR̂_He = blockdiag(R̂^-1) couples the outputs of one time step in H̃ and q̃; with bounds and a missing measurement."""
    import mpc_b200
    N, He = 5, 4
    gm, oms, rng = make(N, 41, nd=0)
    spd = lambda n, s: (lambda Q: s * (Q @ Q.T) / n + 0.3 * s * np.eye(n))(rng.standard_normal((n, n)))
    cov = dict(Rhat=spd(2, 1.0), Qhat=spd(3, 0.2), P0hat=spd(3, 0.5))
    kw = dict(xhatmin=[-0.8] * 3, xhatmax=[0.8] * 3, whatmin=[-0.4] * 3, whatmax=[0.4] * 3, vhatmin=[-3.0] * 2, vhatmax=[3.0] * 2)
    g = mpc_b200.MovingHorizonEstimator(gm, He=He, nint_ym=[0, 0], **cov).setconstraint(**kw)
    os_ = [OMHE(m, He=He, nint_ym=0, **cov).setconstraint(**kw) for m in oms]
    worst, nact = 0.0, 0
    for k in range(2 * He + 3):
        y = np.array([5.0, 3.0]) + rng.standard_normal((N, 2))
        u = np.array([1.0, -2.0]) + rng.standard_normal((N, 2))
        if k == 5:
            y[2, 1] = np.nan
        xg = g.preparestate(y)
        for i, o in enumerate(os_):
            xo = o.preparestate(y[i])
            assert g.status[i] == o.last_qp["status"], (k, i)
            if g.status[i] == 0:
                tol = 2e-6 if g.iters[i] > 0 else 1e-9
                e = np.abs(xg[i] - xo).max() / (1 + np.abs(xo).max())
                ej = abs(g.J[i] - o.Jval) / (1 + abs(o.Jval))
                assert e < tol and ej < 1e-8, (k, i, e, ej, g.iters[i])
                worst = max(worst, e)
                nact += int(g.iters[i] > 0)
            o.updatestate(u[i], y[i])
        g.updatestate(u, y)
    assert nact > 5
    print("MHE dense covariances worst", worst, "active solves", nact)


@pytest.mark.parametrize("direct", [True, False])
def test_mhe_multiple_shooting_layout_matches_ms_oracle(direct):
    """MovingHorizonEstimator(...; transcription=MultipleShooting()) (test/2_test_state_estim.jl:1126-1139): the handle
    solves the condensed problem, the host mirror returns Z̃ = [x̂0(k-Nk+p); X̂0; Ŵ] in the reference's MultipleShooting
    layout.  Checked against oracle/mhe_ms.py, which solves the equality-constrained problem (defect constraints) by a
    generic null-space method: estimate, the whole Z̃ (stage states included, unused entries zero) and J, over the
    growing and the moving window, with state / noise bounds active."""
    import mpc_b200
    from oracle.mhe_ms import MovingHorizonEstimatorMS as OMS
    N, He = 5, 4
    gm, oms, rng = make(N, 6, nd=1)
    # soft bounds: every window stays feasible.  (On an INFEASIBLE window the reference's fallback differs by transcription
    # -- MultipleShooting keeps the shifted stage states of the previous solution, transcription.jl:1037-1076 --, while the
    # handle always returns the open-loop rebuild of the SingleShooting fallback: documented divergence)
    kw = dict(xhatmin=[-0.6] * 3, xhatmax=[0.6] * 3, whatmin=[-0.3] * 3, whatmax=[0.3] * 3,
              vhatmin=[-2.5] * 2, vhatmax=[2.5] * 2, c_xhatmin=[1] * 3, c_xhatmax=[1] * 3, c_whatmin=[1] * 3,
              c_whatmax=[1] * 3, c_vhatmin=[1] * 2, c_vhatmax=[1] * 2)
    g = mpc_b200.MovingHorizonEstimator(gm, He=He, nint_ym=[0, 0], direct=direct, Cwt=1e5,
                                        transcription="MultipleShooting").setconstraint(**kw)
    os_ = [OMS(m, He=He, nint_ym=0, direct=direct, Cwt=1e5).setconstraint(**kw) for m in oms]
    assert g.Ztilde.shape == (N, 1 + 3 + 2 * 3 * He) and os_[0].Ztilde.shape == (1 + 3 + 2 * 3 * He,)
    worst, nact = run_both(g, os_, rng, 2 * He + 3, 1, 2, 2, tol_active=2e-6)
    assert nact > 10
    print("MHE MultipleShooting worst", worst, "active solves", nact)
