"""Pins the numpy MHE oracle (oracle/mhe.py) to the reference's own assertions for the linear
MovingHorizonEstimator."""
import numpy as np
import pytest

from oracle import qp
from oracle.linmpc import LinModel, SteadyKalmanFilter
from oracle.mhe import KalmanFilter, MovingHorizonEstimator


def plant(seed=0, nx=4, nu=2, ny=2, nd=1):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((nx, nx))
    A *= 0.8 / np.abs(np.linalg.eigvals(A)).max()
    return LinModel(A, rng.standard_normal((nx, nu)), rng.standard_normal((ny, nx)), Bd=rng.standard_normal((nx, nd)),
                    Dd=0.1 * rng.standard_normal((ny, nd)), uop=[10, 50], yop=[50, 30], dop=[20]), rng


@pytest.mark.parametrize("He", [3, 5])
def test_mhe_equals_kalman_filter_predictor_form(He):
    # test/2_test_state_estim.jl:1750-1766: direct=false, nint_ym=0, atol=rtol=1e-6
    model, rng = plant(1)
    kf = KalmanFilter(model, nint_ym=0, direct=False)
    mhe = MovingHorizonEstimator(model, He=He, nint_ym=0, direct=False)
    X_mhe, X_kf = [], []
    for i in range(2 * He + 3):
        y = np.array([50, 31]) + rng.standard_normal(2)
        X_mhe.append(mhe.preparestate(y, [25]).copy())
        X_kf.append(kf.preparestate(y, [25]).copy())
        mhe.updatestate([11, 50], y, [25])
        kf.updatestate([11, 50], y, [25])
    assert np.allclose(X_mhe, X_kf, atol=1e-6, rtol=1e-6)
    assert np.abs(np.array(X_mhe) - np.array(X_kf)).max() < 1e-9


def test_mhe_equals_kalman_filter_current_form():
    # test/2_test_state_estim.jl:1767-1784: direct=true, sigmaP_0 recovered from the KF
    model, rng = plant(2)
    kf = KalmanFilter(model, nint_ym=0, direct=True)
    kf.preparestate([50, 30], [20])
    sP = np.sqrt(np.diag(kf.Phat))
    # the corrected covariance is not diagonal in general: start both from the same (diagonalised) P
    kf.Phat = np.diag(sP ** 2)
    mhe = MovingHorizonEstimator(model, He=3, nint_ym=0, direct=True, sigmaP_0=sP)
    kf.updatestate([10, 50], [50, 30], [20])
    X_mhe, X_kf = [], []
    for i in range(9):
        y = np.array([50, 31]) + rng.standard_normal(2)
        X_mhe.append(mhe.preparestate(y, [25]).copy())
        X_kf.append(kf.preparestate(y, [25]).copy())
        mhe.updatestate([11, 50], y, [25])
        kf.updatestate([11, 50], y, [25])
    assert np.allclose(X_mhe, X_kf, atol=1e-6, rtol=1e-6)


def test_mhe_doctest_half():
    # src/estimator/mhe/execute.jl:134-144: A=B=C=1, He=1, direct=false, y=1 -> Yhat = 0.5
    mhe = MovingHorizonEstimator(LinModel([[1.0]], [[1.0]], [[1.0]], Ts=5.0), He=1, nint_ym=0, direct=False)
    mhe.updatestate([0], [1])
    xarr = mhe.Ztilde[:1]
    assert np.round(mhe.Cmhat @ xarr, 3)[0] == 0.5


@pytest.mark.parametrize("Cwt", [1e5, np.inf])
def test_mhe_constraint_violation(Cwt):
    # test/2_test_state_estim.jl:1491-1553 (He=1, nint_ym=0), atol 5e-2
    rng = np.random.default_rng(3)
    A = np.diag([0.8, 0.9])
    model = LinModel(A, np.eye(2) * 0.5, np.eye(2), uop=[10, 50], yop=[50, 30])
    mhe = MovingHorizonEstimator(model, He=1, nint_ym=0, Cwt=Cwt)
    mhe.setconstraint(xhatmin=[-100, -100], xhatmax=[100, 100], whatmin=[-100, -100], whatmax=[100, 100],
                      vhatmin=[-100, -100], vhatmax=[100, 100])
    if np.isfinite(Cwt):
        mhe.setconstraint(c_xhatmin=[1, 1], c_xhatmax=[1, 1], c_whatmin=[0.1, 0.1], c_whatmax=[0.1, 0.1],
                          c_vhatmin=[1, 1], c_vhatmax=[1, 1])

    def step():
        mhe.preparestate([50, 30])
        x = mhe.updatestate([10, 50], [50, 30])
        assert mhe.last_qp["status"] == qp.OPTIMAL
        return x
    big = dict(xhatmin=[-100, -100], xhatmax=[100, 100], whatmin=[-100, -100], whatmax=[100, 100],
               vhatmin=[-100, -100], vhatmax=[100, 100])
    mhe.setconstraint(xhatmin=[1, 1], xhatmax=[100, 100])
    assert step() == pytest.approx([1, 1], abs=5e-2)
    mhe.setconstraint(xhatmin=[-100, -100], xhatmax=[-1, -1])
    assert step() == pytest.approx([-1, -1], abs=5e-2)
    mhe.setconstraint(**big)
    mhe.setconstraint(whatmin=[1, 1], whatmax=[100, 100])
    step()
    assert mhe.Ztilde[-2:] == pytest.approx([1, 1], abs=5e-2)
    mhe.setconstraint(whatmin=[-100, -100], whatmax=[-1, -1])
    step()
    assert mhe.Ztilde[-2:] == pytest.approx([-1, -1], abs=5e-2)
    mhe.setconstraint(**big)
    mhe.setconstraint(vhatmin=[1, 1], vhatmax=[100, 100])
    step()
    assert mhe.Vhat == pytest.approx([1, 1], abs=5e-2)
    mhe.setconstraint(vhatmin=[-100, -100], vhatmax=[-1, -1])
    step()
    assert mhe.Vhat == pytest.approx([-1, -1], abs=5e-2)


def test_host_kalman_filter_mirror_equals_oracle():
    """The batched host-side KalmanFilter of the Python mirror (modelpredictivecontrol.jl_b200/host.py) against the
    oracle's restatement of src/estimator/kalman.jl:1235-1290 (correct + predict, Hermitian(:L)): x̂ and P̂ to 1e-12."""
    import mpc_b200
    from oracle.linmpc import LinModel as OLinModel
    from oracle.mhe import KalmanFilter as OKF
    N = 5
    rng = np.random.default_rng(8)
    A = 0.6 * rng.standard_normal((N, 3, 3)) / 2
    Bu, C = rng.standard_normal((N, 3, 2)), rng.standard_normal((N, 2, 3))
    Bd, Dd = rng.standard_normal((N, 3, 1)), 0.1 * rng.standard_normal((N, 2, 1))
    op = dict(uop=[1.0, -2.0], yop=[5.0, 3.0], dop=[0.5])
    g = mpc_b200.KalmanFilter(mpc_b200.LinModel(A, Bu, C, Bd=Bd, Dd=Dd, N=N, **op), sigmaR=[0.5, 2.0])
    os_ = [OKF(OLinModel(A[i], Bu[i], C[i], Bd=Bd[i], Dd=Dd[i], **op), sigmaR=[0.5, 2.0]) for i in range(N)]
    for k in range(12):
        y = 5 + rng.standard_normal((N, 2)); d = 0.5 + rng.standard_normal((N, 1)); u = rng.standard_normal((N, 2))
        xg = g.preparestate(y, d)
        for i, o in enumerate(os_):
            xo = o.preparestate(y[i], d[i])
            assert np.abs(xg[i] - xo).max() < 1e-12 * (1 + np.abs(xo).max())
            assert np.abs(g.Phat[i] - o.Phat).max() < 1e-12
            o.updatestate(u[i], y[i], d[i])
        g.updatestate(u, y, d)
        assert np.abs(g.Phat - np.stack([o.Phat for o in os_])).max() < 1e-12


def _setup_linmodel_tustin_d():
    """SetupMPCtests' `sys` as LinModel(sys, Ts, i_u=[1,2], i_d=[3]) builds it (src/model/linmodel.jl:165-198): zero-order
    hold for the manipulated inputs, Tustin for the measured disturbance (four states, so the default estimators have
    nx̂ = 6 as the reference's assertions expect), with the operating points of test/2_test_state_estim.jl:1037."""
    from oracle.linmpc import zoh_first_order
    Ts = 400.0
    a1, b1, c1 = (m[0, 0] for m in zoh_first_order(1.90, 1800.0, Ts))
    a2, b2, c2 = (m[0, 0] for m in zoh_first_order(0.74, 800.0, Ts))

    def tustin(k, tau):
        al = Ts / (2 * tau)
        p, g = (1 - al) / (1 + al), k * al / (1 + al)
        return p, g * (1 + p), g
    p1, cd1, dd1 = tustin(1.90, 1800.0)
    p2, cd2, dd2 = tustin(-0.74, 800.0)
    return LinModel(np.diag([a1, a2, p1, p2]), np.array([[b1, b1], [-b2, b2], [0, 0], [0, 0]]),
                    np.array([[c1, 0, cd1, 0], [0, c2, 0, cd2]]), Bd=np.array([[0.0], [0.0], [1.0], [1.0]]),
                    Dd=np.array([[dd1], [dd2]]), Ts=Ts, uop=[10, 50], yop=[50, 30], dop=[5])


@pytest.mark.parametrize("kw,tol", [(dict(), 1e-3), (dict(nint_u=[1, 1], nint_ym=[0, 0], direct=False), 1e-2)])
def test_mhe_estimation_known_answers(kw, tol):
    """test/2_test_state_estim.jl:1039-1078 ("MHE estimation and getinfo (LinModel)", SingleShooting): at the operating
    point the estimate stays zero (atol 1e-9); with a constant input offset, then a constant output offset, the estimated
    output returns to the measurement within 40 periods (atol 1e-3 in the current form, 1e-2 in the prediction form)."""
    mhe = MovingHorizonEstimator(_setup_linmodel_tustin_d(), He=2, **kw)
    assert mhe.nxhat == 6
    mhe.preparestate([50, 30], [5])
    xhat = mhe.updatestate([10, 50], [50, 30], [5])
    assert xhat == pytest.approx(np.zeros(6), abs=1e-9) and mhe.xhat0 == pytest.approx(np.zeros(6), abs=1e-9)
    for u, ym in (([11, 52], [50, 30]), ([10, 50], [51, 32])):
        for _ in range(40):
            mhe.preparestate(ym, [5])
            mhe.updatestate(u, ym, [5])
        if kw.get("direct", True):
            mhe.preparestate(ym, [5])
        assert mhe.evaloutput([5]) == pytest.approx(ym, abs=tol)


def test_mhe_covariance_constructor_and_nan_measurement():
    """test/2_test_state_estim.jl:1080-1093 (full Q̂ / R̂ matrices and P̂_0 taken from a SteadyKalmanFilter, no integrators:
    the estimate stays zero at the operating point) and :1109-1118 (a NaN measurement is left out of the objective,
    mhe/execute.jl:436-441: the estimate at the operating point is still zero)."""
    model = _setup_linmodel_tustin_d()
    Qhat, Rhat = np.diag(np.full(4, 0.25) ** 2), np.eye(2)
    skf = SteadyKalmanFilter(model, sigmaQ=np.full(4, 0.25), sigmaR=[1, 1], nint_u=0, nint_ym=0)
    mhe3 = MovingHorizonEstimator(model, He=2, nint_u=0, nint_ym=0, P0hat=skf.Phat, Qhat=Qhat, Rhat=Rhat)
    mhe3.preparestate([50, 30], [5])
    xhat = mhe3.updatestate([10, 50], [50, 30], [5])
    assert xhat == pytest.approx(np.zeros(4), abs=1e-9) and mhe3.xhat0 == pytest.approx(np.zeros(4), abs=1e-9)
    mhe3.preparestate([50, 30], [5])
    assert mhe3.evaloutput([5]) == pytest.approx([50, 30], abs=1e-9)
    mhe4 = MovingHorizonEstimator(model, He=2, direct=True)
    mhe4.preparestate([50, np.nan], [5])
    assert mhe4.xhat0 == pytest.approx(np.zeros(6), abs=1e-9)
    mhe5 = MovingHorizonEstimator(model, He=2, direct=False)
    mhe5.updatestate([10, 50], [50, np.nan], [5])
    assert np.isfinite(mhe5.xhat0).all()


def test_mhe_setconstraint_known_answers():
    """test/2_test_state_estim.jl:1385-1490 ("MHE set constraints", LinModel parts): per-stage and whole-window bounds and
    softness weights (the first nx̂ entries of X̂min / C_x̂min belong to the arrival state), size checks, and what is frozen
    after the first solve (softness, the +-Inf pattern) or without a slack variable (Cwt = Inf)."""
    from oracle.linmpc import zoh_first_order
    Ts = 400.0
    a1, b1, g1 = (m[0, 0] for m in zoh_first_order(1.90, 1800.0, Ts))
    a2, b2, g2 = (m[0, 0] for m in zoh_first_order(0.74, 800.0, Ts))
    # LinModel(sys, Ts, i_u=[1,2]) with uop = [10, 50], yop = [50, 30]: two states, no measured disturbance
    mk = lambda: LinModel(np.diag([a1, a2]), np.array([[b1, b1], [-b2, b2]]), np.diag([g1, g2]), Ts=Ts, uop=[10, 50], yop=[50, 30])
    mhe1 = MovingHorizonEstimator(mk(), He=1, nint_ym=0, Cwt=1e3)
    c = mhe1.con
    mhe1.setconstraint(xhatmin=[-51, -52], xhatmax=[53, 54])
    assert np.allclose(c["X0min"], [-51, -52]) and np.allclose(c["X0max"], [53, 54])
    assert np.allclose(c["xhat0min"], [-51, -52]) and np.allclose(c["xhat0max"], [53, 54])
    mhe1.setconstraint(whatmin=[-55, -56], whatmax=[57, 58])
    assert np.allclose(c["Wmin"], [-55, -56]) and np.allclose(c["Wmax"], [57, 58])
    mhe1.setconstraint(vhatmin=[-59, -60], vhatmax=[61, 62])
    assert np.allclose(c["Vmin"], [-59, -60]) and np.allclose(c["Vmax"], [61, 62])
    mhe1.setconstraint(c_xhatmin=[0.01, 0.02], c_xhatmax=[0.03, 0.04])
    assert np.allclose(c["C_xmin"], [0.01, 0.02]) and np.allclose(c["C_xmax"], [0.03, 0.04])
    assert np.allclose(c["c_xmin"], [0.01, 0.02]) and np.allclose(c["c_xmax"], [0.03, 0.04])
    mhe1.setconstraint(c_whatmin=[0.05, 0.06], c_whatmax=[0.07, 0.08])
    assert np.allclose(c["C_wmin"], [0.05, 0.06]) and np.allclose(c["C_wmax"], [0.07, 0.08])
    mhe1.setconstraint(c_vhatmin=[0.09, 0.10], c_vhatmax=[0.11, 0.12])
    assert np.allclose(c["C_vmin"], [0.09, 0.10]) and np.allclose(c["C_vmax"], [0.11, 0.12])
    # the softness weights are the first column of every block of the QP's A (construct.jl:962-995): one solve shows them
    mhe1.preparestate([50, 30])
    P = mhe1.build_qp()
    assert np.allclose(-P["A"][:, 0], [0.01, 0.02, 0.03, 0.04, 0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.07, 0.08, 0.09, 0.10, 0.11, 0.12])

    mhe2 = MovingHorizonEstimator(mk(), He=4, nint_ym=0, Cwt=1e3)
    c2 = mhe2.con
    r = lambda lo, hi: np.arange(lo, hi + 1.0)
    mhe2.setconstraint(Xhatmin=-r(1, 10), Xhatmax=r(1, 10))
    assert np.allclose(c2["X0min"], -r(3, 10)) and np.allclose(c2["X0max"], r(3, 10))
    assert np.allclose(c2["xhat0min"], -r(1, 2)) and np.allclose(c2["xhat0max"], r(1, 2))
    mhe2.setconstraint(Whatmin=-r(11, 18), Whatmax=r(11, 18))
    assert np.allclose(c2["Wmin"], -r(11, 18)) and np.allclose(c2["Wmax"], r(11, 18))
    mhe2.setconstraint(Vhatmin=-r(31, 38), Vhatmax=r(31, 38))
    assert np.allclose(c2["Vmin"], -r(31, 38)) and np.allclose(c2["Vmax"], r(31, 38))
    mhe2.setconstraint(C_xhatmin=0.01 * r(1, 10), C_xhatmax=0.02 * r(1, 10))
    assert np.allclose(c2["C_xmin"], 0.01 * r(3, 10)) and np.allclose(c2["C_xmax"], 0.02 * r(3, 10))
    assert np.allclose(c2["c_xmin"], 0.01 * r(1, 2)) and np.allclose(c2["c_xmax"], 0.02 * r(1, 2))
    mhe2.setconstraint(C_whatmin=0.03 * r(11, 18), C_whatmax=0.04 * r(11, 18))
    assert np.allclose(c2["C_wmin"], 0.03 * r(11, 18)) and np.allclose(c2["C_wmax"], 0.04 * r(11, 18))
    mhe2.setconstraint(C_vhatmin=0.05 * r(31, 38), C_vhatmax=0.06 * r(31, 38))
    assert np.allclose(c2["C_vmin"], 0.05 * r(31, 38)) and np.allclose(c2["C_vmax"], 0.06 * r(31, 38))
    for kw in ("xhatmin", "xhatmax", "whatmin", "whatmax", "vhatmin", "vhatmax",
               "c_xhatmin", "c_xhatmax", "c_whatmin", "c_whatmax", "c_vhatmin", "c_vhatmax"):
        with pytest.raises(ValueError):       # DimensionMismatch
            mhe2.setconstraint(**{kw: [1.0]})
    assert np.allclose(c2["C_vmax"], 0.06 * r(31, 38)) and np.allclose(c2["X0min"], -r(3, 10))  # nothing stored by a failed call
    mhe1.updatestate([10, 50], [50, 30])
    inf = np.inf
    for kw, v in (("xhatmin", [-inf, -inf]), ("xhatmax", [inf, inf]), ("whatmin", [-inf, -inf]), ("whatmax", [inf, inf]),
                  ("vhatmin", [-inf, -inf]), ("vhatmax", [inf, inf]), ("c_xhatmin", [100, 100]), ("c_xhatmax", [200, 200]),
                  ("c_whatmin", [300, 300]), ("c_whatmax", [400, 400]), ("c_vhatmin", [500, 500]), ("c_vhatmax", [600, 600])):
        with pytest.raises(RuntimeError):     # ErrorException: frozen after the first solve
            mhe1.setconstraint(**{kw: v})
    mhe1.setconstraint(xhatmin=[-61, -62])    # finite values may still move
    assert np.allclose(c["xhat0min"], [-61, -62])
    mhe4 = MovingHorizonEstimator(mk(), He=1, nint_ym=0, Cwt=inf)
    for kw in ("c_xhatmin", "c_xhatmax", "c_whatmin", "c_whatmax", "c_vhatmin", "c_vhatmax"):
        with pytest.raises(ValueError):       # ArgumentError: no slack variable
            mhe4.setconstraint(**{kw: [1, 1]})


@pytest.mark.parametrize("direct", [True, False])
def test_mhe_unfilled_window_input_disturbance(direct):
    """test/2_test_state_estim.jl:1313-1336 ("MHE estimation with unfilled window"): x(k+1) = 0.5 x + u, y = x (the
    reference builds it as a NonLinModel; the dynamics are linear, so the linear MHE solves the same windows), the plant
    receives u = 0.1 while the estimator is told u = 0: with an input integrator (nint_u = 1) the estimated output equals
    the plant output within 1e-6, from the growing window (He = 3) on."""
    mk = lambda: LinModel(np.array([[0.5]]), np.array([[1.0]]), np.array([[1.0]]), Ts=10.0)
    plant = mk()
    mhe = MovingHorizonEstimator(mk(), He=3, nint_u=[1], direct=direct)
    for _ in range(40):
        y = plant.evaloutput()
        mhe.preparestate(y)
        mhe.updatestate([0.0], y)
        plant.updatestate([0.1])
    mhe.preparestate(plant.evaloutput())
    assert mhe.evaloutput() == pytest.approx(plant.evaloutput(), abs=1e-6)


def test_mhe_construction_known_answers():
    """test/2_test_state_estim.jl:886-976 ("MHE construction (LinModel)", the parts of the linear SingleShooting path):
    sizes of the augmented state and of the decision vector with and without the slack variable, the measured-output
    selection, covariances built from the standard deviations, window lengths, integrator choices, error cases."""
    model = _setup_linmodel_tustin_d()
    mhe1 = MovingHorizonEstimator(model, He=5)
    assert (mhe1.nym, mhe1.nxhat) == (2, 6) and mhe1.nxhat - model.nx == 2
    assert mhe1.E.shape[1] == 6 * mhe1.nxhat and mhe1.nZ == mhe1.nxhat + mhe1.nxhat * 5
    mhe3 = MovingHorizonEstimator(model, He=5, i_ym=[1])                       # the reference's i_ym=[2] (1-based)
    assert (mhe3.nym, model.ny - mhe3.nym, mhe3.nxhat) == (1, 1, 5)
    mhe4 = MovingHorizonEstimator(model, He=5, sigmaQ=[1, 2, 3, 4], sigmaQint_ym=[5, 6], sigmaR=[7, 8])
    assert np.array_equal(mhe4.Qhat, np.diag([1.0, 4, 9, 16, 25, 36])) and np.array_equal(mhe4.Rhat, np.diag([49.0, 64]))
    assert MovingHorizonEstimator(model, He=5, nint_ym=[2, 2]).nxhat == 8
    mhe6 = MovingHorizonEstimator(model, He=5, sigmaP_0=[1, 2, 3, 4], sigmaPint_ym_0=[5, 6])
    assert np.array_equal(mhe6.P0hat, np.diag([1.0, 4, 9, 16, 25, 36])) and np.array_equal(mhe6.Parr_old, mhe6.P0hat)
    assert mhe6.Parr_old is not mhe6.P0hat
    assert np.allclose(mhe6.invPbar, np.linalg.inv(np.diag(np.arange(1, 7.0) ** 2)))
    mhe7 = MovingHorizonEstimator(model, He=10)
    assert (mhe7.X0_old.size, mhe7.Y0m.size, mhe7.U0.size, mhe7.D0.size) == (60, 20, 20, 11)
    mhe8 = MovingHorizonEstimator(model, He=5, nint_u=[1, 1], nint_ym=[0, 0])
    assert mhe8.nxhat == 6 and list(mhe8.nint_u) == [1, 1] and list(mhe8.nint_ym) == [0, 0]
    mhe12 = MovingHorizonEstimator(model, He=5, Cwt=1e3)
    assert mhe12.nZ == 6 * mhe12.nxhat + 1 and mhe12.Cwt == 1e3
    for kw in (dict(He=0), dict(He=5, Cwt=-1)):
        with pytest.raises(ValueError):
            MovingHorizonEstimator(model, **kw)
    with pytest.raises(TypeError):  # He has no default (ArgumentError in the reference)
        MovingHorizonEstimator(model)


def test_kalman_filter_estimator_methods_and_setmodel():
    """test/2_test_state_estim.jl:207-265 ("KF estimator methods") and :268-294 ("KF set model"): the time-varying
    KalmanFilter the step kernels can run fused (bmpc_set_estimator_cov) and that `setmodel!` needs."""
    from oracle.linmpc import zoh_first_order
    Ts = 400.0
    a1, b1, g1 = (m[0, 0] for m in zoh_first_order(1.90, 1800.0, Ts))
    a2, b2, g2 = (m[0, 0] for m in zoh_first_order(0.74, 800.0, Ts))
    mk = lambda: LinModel(np.diag([a1, a2]), np.array([[b1, b1], [-b2, b2]]), np.diag([g1, g2]), Ts=Ts, uop=[10, 50], yop=[50, 30])
    kf1 = KalmanFilter(mk())
    kf1.preparestate([50, 30])
    assert kf1.updatestate([10, 50], [50, 30]) == pytest.approx(np.zeros(4), abs=1e-12)
    kf1.preparestate([50, 30])
    assert kf1.evaloutput() == pytest.approx([50, 30])
    assert kf1.initstate([10, 50], [50, 30 + 1]) == pytest.approx([0, 0, 0, 1], abs=1e-9)
    kf1.setstate([1, 2, 3, 4], np.diag([0.1, 0.2, 0.3, 0.4]))
    assert kf1.xhat0 == pytest.approx([1, 2, 3, 4]) and np.allclose(kf1.Phat, np.diag([0.1, 0.2, 0.3, 0.4]))
    for est, prep in ((kf1, True), (KalmanFilter(mk(), nint_u=[1, 1], direct=False), False)):
        for uu, ym in (([11, 52], [50, 30]), ([10, 50], [51, 32])):
            for _ in range(40):
                est.preparestate(ym)
                est.updatestate(uu, ym)
            if prep:
                est.preparestate(ym)
            assert est.evaloutput() == pytest.approx(ym, abs=1e-3)
    kf4 = KalmanFilter(mk(), direct=True)
    kf4.xhat0[:] = 7
    kf4.preparestate([55, np.nan])
    assert (kf4.xhat0 == 7).all()
    kf5 = KalmanFilter(mk(), direct=False)
    kf5.updatestate([10, 50], [55, np.nan])
    assert np.isfinite(kf5.xhat0).all()
    # set model
    lin = lambda a, uop, yop, xop: LinModel([[a]], [[0.3]], [[1.0]], Ts=10.0, uop=[uop], yop=[yop], xop=[xop], fop=[xop])
    kf = KalmanFilter(lin(0.5, 2.0, 50.0, 3.0), nint_ym=0)
    assert np.allclose(kf.Ahat, [[0.5]])
    kf.preparestate([50.0])
    assert kf.evaloutput() == pytest.approx([50.0])
    kf.preparestate([50.0])
    assert kf.updatestate([2.0], [50.0]) == pytest.approx([3.0])
    kf.setmodel(lin(0.2, 3.0, 55.0, 3.0))
    assert np.allclose(kf.Ahat, [[0.2]])
    kf.preparestate([55.0])
    assert kf.evaloutput() == pytest.approx([55.0])
    kf.preparestate([55.0])
    assert kf.updatestate([3.0], [55.0]) == pytest.approx([3.0])
    kf.setmodel(lin(0.2, 3.0, 55.0, 8.0))
    assert kf.xhat0 == pytest.approx([3.0 - 8.0])
    kf.setmodel(kf.model, Qhat=[1e-3], Rhat=[1e-6])
    assert np.allclose(kf.Qhat, [[1e-3]]) and np.allclose(kf.Rhat, [[1e-6]])


def test_host_estimator_mirrors_nan_and_initstate():
    """Host mirror (modelpredictivecontrol.jl_b200/host.py, runs on the CPU) against the oracle for the estimator behaviour
    the reference tests in test/2_test_state_estim.jl:77,114-126,220,252-264: `initstate!` and a NaN measurement (the
    instance that has one skips its correction, the others do not), for SteadyKalmanFilter and KalmanFilter."""
    import mpc_b200
    from oracle.linmpc import LinModel as OLinModel
    from oracle.mhe import KalmanFilter as OKF
    N = 4
    rng = np.random.default_rng(3)
    A = 0.5 * rng.standard_normal((N, 3, 3)) / 2
    Bu, C = rng.standard_normal((N, 3, 2)), rng.standard_normal((N, 2, 3))
    op = dict(uop=[1.0, -2.0], yop=[5.0, 3.0])
    for G, O in ((mpc_b200.SteadyKalmanFilter, SteadyKalmanFilter), (mpc_b200.KalmanFilter, OKF)):
        g = G(mpc_b200.LinModel(A, Bu, C, N=N, **op))
        os_ = [O(OLinModel(A[i], Bu[i], C[i], **op)) for i in range(N)]
        u, y = 1 + rng.standard_normal((N, 2)), 5 + rng.standard_normal((N, 2))
        xg = g.initstate(u, y)
        for i, o in enumerate(os_):
            assert np.abs(xg[i] - o.initstate(u[i], y[i])).max() < 1e-9
        for k in range(6):
            y = 5 + rng.standard_normal((N, 2))
            if k == 2:
                y[1, 0] = np.nan          # one instance, one channel
            u = rng.standard_normal((N, 2))
            xg = g.preparestate(y)
            for i, o in enumerate(os_):
                xo = o.preparestate(y[i])
                assert np.isfinite(xg[i]).all() and np.abs(xg[i] - xo).max() < 1e-11 * (1 + np.abs(xo).max()), (G.__name__, k, i)
                if hasattr(o, "Phat") and G is mpc_b200.KalmanFilter:
                    assert np.abs(g.Phat[i] - o.Phat).max() < 1e-11
                o.updatestate(u[i], y[i])
            g.updatestate(u, y)


def test_kalman_filter_construction():
    """test/2_test_state_estim.jl:155-205 ("KF construction"): augmentation sizes, covariances from standard deviations."""
    from oracle.linmpc import zoh_first_order
    Ts = 400.0
    a1, b1, g1 = (m[0, 0] for m in zoh_first_order(1.90, 1800.0, Ts))
    a2, b2, g2 = (m[0, 0] for m in zoh_first_order(0.74, 800.0, Ts))
    linmodel = LinModel(np.diag([a1, a2]), np.array([[b1, b1], [-b2, b2]]), np.diag([g1, g2]), Ts=Ts, uop=[10, 50], yop=[50, 30])
    k1 = KalmanFilter(linmodel)
    assert (len(k1.i_ym), k1.nxs, k1.nxhat, list(k1.nint_ym)) == (2, 2, 4, [1, 1])
    linmodel2 = _setup_linmodel_tustin_d()
    k2 = KalmanFilter(linmodel2, i_ym=[1])
    assert (len(k2.i_ym), k2.nxs, k2.nxhat) == (1, 1, 5)
    assert (KalmanFilter(linmodel, nint_ym=0).nxs, KalmanFilter(linmodel, nint_ym=0).nxhat) == (0, 2)
    assert (KalmanFilter(linmodel, nint_ym=[2, 2]).nxs, KalmanFilter(linmodel, nint_ym=[2, 2]).nxhat) == (4, 6)
    k5 = KalmanFilter(linmodel2, sigmaQ=[1, 2, 3, 4], sigmaQint_ym=[5, 6], sigmaR=[7, 8])
    assert np.array_equal(k5.Qhat, np.diag([1.0, 4, 9, 16, 25, 36])) and np.array_equal(k5.Rhat, np.diag([49.0, 64]))
    k6 = KalmanFilter(linmodel2, sigmaP_0=[1, 2, 3, 4], sigmaPint_ym_0=[5, 6])
    assert np.array_equal(k6.Phat, np.diag([1.0, 4, 9, 16, 25, 36]))
    k7 = KalmanFilter(linmodel, nint_u=[1, 1])
    assert (k7.nxs, k7.nxhat, list(k7.nint_u), list(k7.nint_ym)) == (2, 4, [1, 1], [0, 0])
    with pytest.raises(ValueError):
        KalmanFilter(linmodel, nint_ym=0, sigmaP_0=[1])
