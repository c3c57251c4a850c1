"""Pins the numpy MHE oracle (oracle/mhe.py) to the reference's own assertions for the linear
MovingHorizonEstimator."""
import numpy as np
import pytest

from oracle import qp
from oracle.linmpc import LinModel
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
