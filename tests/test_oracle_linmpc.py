"""Pins the numpy oracle (oracle/linmpc.py, oracle/qp.py) to every inline known answer the
reference's own tests / doctests hold for the LinMPC path.
Each test cites the reference assertion it restates; tolerances are the reference's own."""
import numpy as np
import pytest
import scipy.linalg

from oracle import qp
from oracle.linmpc import (ExplicitMPC, LinModel, LinMPC, ManualEstimator, SteadyKalmanFilter,
                           move_blocking, zoh_first_order)


def first_order(gain, tau, Ts, **op):
    A, B, C = zoh_first_order(gain, tau, Ts)
    return LinModel(A, B, C, Ts=Ts, **op)


def test_move_blocking():
    # docstring example src/controller/construct.jl:624-627 and the Int method :653-660
    assert move_blocking(10, [1, 2, 3, 6, 7]) == [1, 2, 3, 4]
    assert move_blocking(10, [1, 2, 3]) == [1, 2, 3, 4]
    assert move_blocking(20, 5) == [1, 1, 1, 1, 16]


def test_moveinput_known_answer():
    # test/3_test_predictive_control.jl:93-106 and the moveinput! doctest execute.jl:49-57
    linmodel = first_order(5, 2, 3.0, yop=[10])
    mpc1 = LinMPC(linmodel, Nwt=[0], Hp=1000, Hc=1)
    r = [15]
    mpc1.preparestate([10])
    u = mpc1.moveinput(r)
    assert u == pytest.approx([1], abs=1e-2)
    u = mpc1.moveinput(r, lastu=[-1])
    assert u == pytest.approx([1], abs=1e-2)
    info = mpc1.getinfo()
    assert info["u"] == pytest.approx(u)
    assert info["Yhat"][-1] == pytest.approx(r[0], abs=1e-2)
    assert info["DU"] == pytest.approx([2.0], abs=1e-2)
    assert info["J"] == pytest.approx(info["J_quad"], rel=1e-9, abs=1e-9)  # construct.jl:821-833
    mpc2 = LinMPC(linmodel, Nwt=[0], Cwt=np.inf, Hp=1000, Hc=1)
    mpc2.preparestate([10])
    assert mpc2.moveinput(r) == pytest.approx([1], abs=1e-2)
    # :111-114  input-setpoint tracking
    mpc3 = LinMPC(linmodel, Mwt=[0], Nwt=[0], Lwt=[1])
    mpc3.preparestate([10])
    u = mpc3.moveinput([0], Rhat_u=np.full(mpc3.Hp, 12.0))
    assert u == pytest.approx([12], abs=1e-2)


def test_measured_disturbance_feedforward():
    # :128-134  LinModel([tf(5,[2000,1]) tf(7,[8000,1])], 3000, i_d=[2]); d=0.1 -> y->0.7, u ~ 0
    Ts = 3000.0
    a1, b1, c1 = zoh_first_order(5, 2000, Ts)
    a2, b2, c2 = zoh_first_order(7, 8000, Ts)
    model = LinModel(np.diag([a1[0, 0], a2[0, 0]]), [[b1[0, 0]], [0]], [[c1[0, 0], c2[0, 0]]],
                     Bd=[[0], [b2[0, 0]]], Dd=[[0]], Ts=Ts)
    mpc6 = LinMPC(model, Nwt=[0], Hp=1000, Hc=1)
    mpc6.preparestate([0], [0])
    d = np.array([0.1])
    u = mpc6.moveinput(7 * d, d)
    assert u == pytest.approx([0], abs=1e-2)
    # :142-150 infeasible problem -> error status -> shifted last solution (zeros) is returned
    mpc_inf = LinMPC(model, Hp=1, Hc=1, Cwt=np.inf).setconstraint(umin=[+1], umax=[-1])
    mpc_inf.preparestate([0], [0])
    u = mpc_inf.moveinput([0], [0])
    assert mpc_inf.last_status == qp.INFEASIBLE
    assert u == pytest.approx([0.0])


def test_move_blocking_zeros():
    # :135-140  Hc=[1,2,3,4], Nwt=10: the held moves are exactly zero
    linmodel = first_order(5, 2, 3.0, yop=[10])
    mpc7 = LinMPC(linmodel, Hp=10, Hc=[1, 2, 3, 4], Nwt=[10])
    mpc7.preparestate([10])
    mpc7.moveinput([15])
    dU = np.diff(mpc7.getinfo()["U"])
    assert dU[[1, 3, 4, 6, 7, 8]] == pytest.approx(np.zeros(6), abs=1e-9)
    assert np.abs(dU[[0, 2, 5]]).min() > 1e-6


def test_manual_estimator_equals_default():
    # :211-237  atol 1e-9
    linmodel = first_order(5, 2, 3.0, yop=[10])
    plant = first_order(5, 2, 3.0, yop=[10])
    r, outdist = [15], np.array([5.0])
    mpc_man = LinMPC(ManualEstimator(linmodel))
    skf = SteadyKalmanFilter(linmodel)
    mpc_def = LinMPC(linmodel)
    U_man, U_def = np.zeros(25), np.zeros(25)
    for i in range(25):
        ym = plant.evaloutput() - outdist
        xhat = skf.preparestate(ym)
        mpc_man.setstate(xhat)
        mpc_def.preparestate(ym)
        u_man, u_def = mpc_man.moveinput(r), mpc_def.moveinput(r)
        U_man[i], U_def[i] = u_man[0], u_def[0]
        skf.updatestate(u_man, ym)
        mpc_def.updatestate(u_def, ym)
        plant.updatestate(u_man)
    assert U_man == pytest.approx(U_def, abs=1e-9)
    # closed loop reaches the setpoint despite the output disturbance (integral action)
    assert plant.evaloutput()[0] - outdist[0] == pytest.approx(15, abs=1e-1)


def test_golden_doctest_17_577311():
    # ext/LinearMPCext.jl:252-261: LinMPC(LinModel(tf(2,[10,1]),1.0)); preparestate!(mpc,[1.0]);
    # moveinput!(mpc,[10.0]) -> 17.577311.  Realisation xdot=-0.1x+0.5u, y=0.4x.
    A, B, C = zoh_first_order(2, 10, 1.0, b=0.5)
    mpc = LinMPC(LinModel(A, B, C, Ts=1.0))
    assert (mpc.Hp, mpc.Hc, mpc.estim.nxhat) == (10, 2, 2)
    mpc.preparestate([1.0])
    u = mpc.moveinput([10.0])
    assert round(float(u[0]), 6) == 17.577311


def test_lqr_equivalence():
    # test/3_test_predictive_control.jl:498-527  atol 1e-5 (terminal cost = DARE solution)
    A = np.array([[0.5, -0.4], [0.6, 0.5]])
    B = C = np.eye(2)
    model, plant = LinModel(A, B, C), LinModel(A, B, C)
    Q, R = np.eye(2), 0.5 * np.eye(2)
    P = scipy.linalg.solve_discrete_are(A, B, Q, R)
    K = np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
    M_Hp = np.block([[np.eye(4), np.zeros((4, 2))], [np.zeros((2, 4)), P]])
    mpc = LinMPC(model, Hp=3, Hc=3, M_Hp=M_Hp, Nwt=[0, 0], Lwt=[0.5, 0.5], nint_ym=0)
    mpc.setstate([1, 1])
    plant.setstate([1, 1])
    X_mpc, X_lqr = np.zeros((2, 20)), np.zeros((2, 20))
    for i in range(20):
        y = plant.evaloutput()
        mpc.preparestate(y)
        u = mpc.moveinput([0, 0])
        X_mpc[:, i] = plant.x0
        mpc.updatestate(u, y)
        plant.updatestate(u)
    x = np.array([1.0, 1.0])
    for i in range(20):
        X_lqr[:, i] = x
        x = A @ x + B @ (-K @ x)
    assert np.abs(X_mpc - X_lqr).max() < 1e-5


@pytest.mark.parametrize("Cwt", [1e5, np.inf])
def test_constraint_violation(Cwt):
    # test/3_test_predictive_control.jl:391-464 (test_bound_violation, soft then hard), atol 1e-1
    A, B, C = zoh_first_order(2, 10, 3.0)
    mpc = LinMPC(LinModel(A, B, C, Ts=3.0), Hp=50, Hc=5, Cwt=Cwt)
    mpc.setconstraint(xhatmin=[-1e6, -np.inf], xhatmax=[1e6, np.inf])
    mpc.setconstraint(umin=[-10], umax=[10])
    mpc.setconstraint(dumin=[-15], dumax=[15])
    mpc.setconstraint(ymin=[-100], ymax=[100])
    if np.isfinite(Cwt):
        mpc.setconstraint(c_xhatmin=[1, 1], c_xhatmax=[1, 1])
        mpc.setconstraint(c_umin=[0.1], c_umax=[0.1])
        mpc.setconstraint(c_dumin=[0.1], c_dumax=[0.1])
        mpc.setconstraint(c_ymin=[1], c_ymax=[1])
    mpc.preparestate([0])
    tol = 1e-1

    def info_after(r):
        mpc.moveinput(r)
        assert mpc.last_status == qp.OPTIMAL and mpc.last_qp["kkt"] < 1e-8
        return mpc.getinfo()
    mpc.setconstraint(umin=[-3], umax=[4])
    assert np.allclose(info_after([-100])["U"], -3, atol=tol)
    assert np.allclose(info_after([100])["U"], 4, atol=tol)
    mpc.setconstraint(umin=[-10], umax=[10])
    mpc.setconstraint(dumin=[-1.5], dumax=[1.25])
    assert np.allclose(info_after([-100])["DU"], -1.5, atol=tol)
    assert np.allclose(info_after([100])["DU"], 1.25, atol=tol)
    mpc.setconstraint(dumin=[-15], dumax=[15])
    mpc.setconstraint(ymin=[-0.5], ymax=[0.9])
    assert np.allclose(info_after([-100])["Yhat"], -0.5, atol=tol)
    assert np.allclose(info_after([100])["Yhat"], 0.9, atol=tol)
    mpc.setconstraint(ymin=[-100], ymax=[100])
    mpc.setconstraint(Ymin=np.r_[-0.5, np.full(49, -100.0)], Ymax=np.r_[0.9, np.full(49, 100.0)])
    info = info_after([-10])
    assert info["Yhat"][0] == pytest.approx(-0.5, abs=tol) and info["Yhat"][-1] == pytest.approx(-10, abs=tol)
    info = info_after([10])
    assert info["Yhat"][0] == pytest.approx(0.9, abs=tol) and info["Yhat"][-1] == pytest.approx(10, abs=tol)
    mpc.setconstraint(ymin=[-100], ymax=[100])
    mpc.setconstraint(xhatmin=[-1e-6, -np.inf], xhatmax=[1e-6, np.inf])
    assert info_after([-100])["xhatend"][0] == pytest.approx(0, abs=tol)
    assert info_after([100])["xhatend"][0] == pytest.approx(0, abs=tol)
    mpc.setconstraint(xhatmin=[-1e6, -np.inf], xhatmax=[1e6, np.inf])
    # construct.jl:548-551
    with pytest.raises(RuntimeError):
        mpc.setconstraint(umin=[-np.inf])


def test_explicit_equals_linmpc_unconstrained():
    # test/3_test_predictive_control.jl:1593-1634 (ExplicitMPC == LinMPC), explicitmpc.jl:209
    rng = np.random.default_rng(0)
    A = np.diag([0.9, 0.7, 0.5]) + 0.05 * rng.standard_normal((3, 3))
    model = LinModel(A, rng.standard_normal((3, 2)), rng.standard_normal((2, 3)))
    lin, exp = LinMPC(model, Hp=12, Hc=[1, 2, 3]), ExplicitMPC(model, Hp=12, Hc=[1, 2, 3])
    for mpc in (lin, exp):
        mpc.preparestate([0.3, -0.2])
    u1, u2 = lin.moveinput([1, -1]), exp.moveinput([1, -1])
    assert u1 == pytest.approx(u2, rel=1e-9, abs=1e-9)
    assert lin.Ztilde[-1] == pytest.approx(0.0, abs=1e-12)


def test_qp_solver_against_bruteforce():
    # exactness of oracle/qp.py itself: enumerate active sets of small random QPs
    import itertools
    rng = np.random.default_rng(1)
    for trial in range(20):
        n, m = 3, 5
        M = rng.standard_normal((n, n))
        H = M @ M.T + 0.1 * np.eye(n)
        q = rng.standard_normal(n) * 3
        G = rng.standard_normal((m, n))
        h = rng.random(m) + 0.1
        best, bz = np.inf, None
        for k in range(0, n + 1):
            for act in itertools.combinations(range(m), k):
                act = list(act)
                if act:
                    Ga = G[act]
                    KKT = np.block([[H, Ga.T], [Ga, np.zeros((k, k))]])
                    try:
                        sol = np.linalg.solve(KKT, np.r_[-q, h[act]])
                    except np.linalg.LinAlgError:
                        continue
                    z, lam = sol[:n], sol[n:]
                    if (lam < -1e-12).any():
                        continue
                else:
                    z = np.linalg.solve(H, -q)
                if (G @ z - h > 1e-10).any():
                    continue
                J = 0.5 * z @ H @ z + q @ z
                if J < best:
                    best, bz = J, z
        sol = qp.solve_qp(H, q, G, h)
        assert sol["status"] == qp.OPTIMAL
        assert sol["z"] == pytest.approx(bz, abs=1e-9)


def _model2_with_disturbance():
    # test/3_test_predictive_control.jl:466-467: LinModel([tf(2,[10,1]) tf(0.1,[7,1])], 3.0, i_d=[2]) with
    # uop=25, dop=30, yop=50, realised as two first-order ZOH states (SURVEY App. D-3 for the (b, c) split)
    A1, B1, C1 = zoh_first_order(2, 10, 3.0)
    A2, B2, C2 = zoh_first_order(0.1, 7, 3.0)
    A = np.diag([A1[0, 0], A2[0, 0]])
    Bu, Bd = np.array([[B1[0, 0]], [0.0]]), np.array([[0.0], [B2[0, 0]]])
    C = np.array([[C1[0, 0], C2[0, 0]]])
    return LinModel(A, Bu, C, Bd=Bd, Dd=np.zeros((1, 1)), Ts=3.0, uop=[25], dop=[30], yop=[50])


@pytest.mark.parametrize("kw,wmin,wmax,steps", [
    (dict(Wy=[[1]]), 36, 75, [(0, "Yhat", 36), (100, "Yhat", 75)]),
    (dict(Wu=[[1]]), 4, 20, [(0, "U", 4), (100, "U", 20)]),
    (dict(Wd=[[1]], Wy=[[1]]), 56, 95, [(0, "Yhat", 56 - 30), (100, "Yhat", 95 - 30)]),
    (dict(Wr=[[1]], Wy=[[1]]), 52, 175, [(21, "Yhat", 52 - 21), (100, "Yhat", 175 - 100)]),
])
def test_custom_linear_constraints_known_answers(kw, wmin, wmax, steps):
    """test/3_test_predictive_control.jl:468-496: custom linear constraints Wy/Wu/Wd/Wr (hard, Cwt = Inf): the whole
    predicted trajectory sits on the active custom bound (the reference's own atol = 1e-1)."""
    mpc = LinMPC(_model2_with_disturbance(), Nwt=[0], Cwt=np.inf, Hp=50, Hc=50, **kw)
    mpc.setconstraint(wmin=[wmin], wmax=[wmax])
    mpc.preparestate([50], [30])
    for ry, key, expect in steps:
        mpc.moveinput([ry], [30])
        assert mpc.last_status == qp.OPTIMAL
        assert np.allclose(mpc.getinfo()[key], expect, atol=1e-1), (kw, ry, mpc.getinfo()[key][:5])


def test_custom_linear_constraints_matrices():
    """test/3_test_predictive_control.jl:53-64: W̄y = repeatdiag(Wy, Hp+1) and friends; relaxW's Ew (construct.jl:1138-1160)
    reproduces W = Ew Z + Fw for a random Z."""
    rng = np.random.default_rng(0)
    m = _model2_with_disturbance()
    Wy, Wu, Wd, Wr = rng.standard_normal((2, 1)), rng.standard_normal((2, 1)), rng.standard_normal((2, 1)), rng.standard_normal((2, 1))
    mpc = LinMPC(m, Hp=7, Hc=3, Wy=Wy, Wu=Wu, Wd=Wd, Wr=Wr)
    assert np.allclose(mpc.Wbar_y, np.kron(np.eye(8), Wy)) and mpc.Wbar_y.shape == (16, 8)
    mpc.preparestate([52], [31])
    mpc.setconstraint(wmin=[-1e3, -1e3], wmax=[1e3, 1e3])
    mpc.moveinput([55], [31])
    Z = mpc.Ztilde
    info = mpc.getinfo()
    Ye = np.concatenate([info["yhat"], info["Yhat"]])
    Ue = np.concatenate([info["U"], info["U"][-1:]])
    De = np.full(8, 31.0)
    Re = np.full(8, 55.0)
    W = mpc.Wbar_y @ Ye + mpc.Wbar_u @ Ue + mpc.Wbar_d @ De + mpc.Wbar_r @ Re
    assert np.allclose(W, mpc.Ew @ Z[:3] + mpc.con.Fw, atol=1e-9)


def test_setmodel_known_answers():
    """test/3_test_predictive_control.jl:529-568 ("LinMPC set model"): operating-point change re-expresses the bounds and
    u0(k-1), a gain change moves the steady-state input (u ~ 2, 15, 13), weights can be replaced."""
    from oracle.mhe import KalmanFilter
    A, B, C = zoh_first_order(5, 2, 3.0)
    mk = lambda gain, yop, uop: LinModel(*zoh_first_order(gain, 2, 3.0), Ts=3.0, yop=[yop], uop=[uop])
    mpc = LinMPC(KalmanFilter(mk(5, 10, 1)), Nwt=[0], Cwt=1e4, Hp=1000, Hc=1)
    mpc.setconstraint(umin=[-24], umax=[26])
    mpc.setconstraint(ymin=[-54], ymax=[56])
    assert np.allclose(mpc.Yop, 10) and np.allclose(mpc.Uop, 1)
    assert np.allclose(mpc.con.U0min, -25) and np.allclose(mpc.con.U0max, 25)
    assert np.allclose(mpc.con.Y0min, -64) and np.allclose(mpc.con.Y0max, 46)
    mpc.preparestate([10])
    u = mpc.moveinput([15])
    assert u == pytest.approx([2], abs=1e-2) and mpc.lastu0 == pytest.approx([1], abs=1e-2)
    mpc.setmodel(mk(5, 20, 11))
    assert np.allclose(mpc.Yop, 20) and np.allclose(mpc.Uop, 11)
    assert np.allclose(mpc.con.U0min, -24.0 - 11) and np.allclose(mpc.con.U0max, 26.0 - 11)
    assert np.allclose(mpc.con.Y0min, -54.0 - 20) and np.allclose(mpc.con.Y0max, 56.0 - 20)
    assert mpc.lastu0 == pytest.approx([2 - 11], abs=1e-2)
    u = mpc.moveinput([40])
    assert u == pytest.approx([15], abs=1e-2)
    mpc.setmodel(mk(10, 20, 11))
    u = mpc.moveinput([40])
    assert u == pytest.approx([13], abs=1e-2)
    mpc.setmodel(Mwt=[100], Nwt=[200], Lwt=[300])
    assert np.allclose(mpc.M_Hp, np.diag(np.full(1000, 100.0)))
    assert np.allclose(mpc.Ntilde_Hc, np.diag([200.0, 1e4]))
    assert np.allclose(mpc.L_Hp, np.diag(np.full(1000, 300.0)))
    mpc.setmodel(M_Hp=np.diag(np.arange(1.0, 1001)), Ntilde_Hc=np.diag([0.1, 1e6]), L_Hp=np.diag(np.arange(1.1, 1001)))
    assert np.allclose(mpc.M_Hp, np.diag(np.arange(1.0, 1001))) and np.allclose(mpc.Ntilde_Hc, np.diag([0.1, 1e6]))
    with pytest.raises(RuntimeError):
        LinMPC(mk(5, 10, 1)).setmodel(mk(5, 20, 11))  # SteadyKalmanFilter: kalman.jl:229-234


def test_internalmodel_step_disturbance_rejection():
    """test/3_test_predictive_control.jl:159-176 ("LinMPC step disturbance rejection", InternalModel): with a constant
    output disturbance of -5 the loop settles at ym = r = 15 with u = 2 (plant gain 5, yop = 10: 10 + 5*2 - 5 = 15)."""
    from oracle.linmpc import InternalModel
    mk = lambda: LinModel(*zoh_first_order(5, 2, 3.0), Ts=3.0, yop=[10])
    plant = mk()
    mpc = LinMPC(InternalModel(mk()))
    u = ym = None
    for i in range(25):
        ym = plant.evaloutput() - 5
        mpc.preparestate(ym)
        u = mpc.moveinput([15])
        mpc.updatestate(u, ym)
        plant.updatestate(u)
    assert u == pytest.approx([2], abs=1e-2) and ym == pytest.approx([15], abs=1e-2)


def test_internalmodel_estimator_methods():
    """test/2_test_state_estim.jl:475-513 (IM estimator methods): unit offsets land in x̂s, ŷ reproduces the measurement."""
    from oracle.linmpc import InternalModel
    rng = np.random.default_rng(2)
    A = np.diag([0.5, 0.7])
    m = LinModel(A, np.eye(2), np.eye(2), uop=[10, 50], yop=[50, 30])
    im = InternalModel(m)
    u, y = [10, 50], np.array([51.0, 31.0])
    im.preparestate(y)
    assert im.updatestate(u, y) == pytest.approx(np.zeros(2))
    im.preparestate(y)
    assert im.updatestate(u, y) == pytest.approx(np.zeros(2))
    assert im.xs == pytest.approx(np.ones(2))
    im.preparestate(y)
    assert im.evaloutput() == pytest.approx([51, 31])
    assert im.initstate([10, 50], [50, 30]) == pytest.approx(np.zeros(2)) and im.xs == pytest.approx(np.zeros(2))
    with pytest.raises(ValueError):
        InternalModel(LinModel(np.eye(1), np.ones((1, 1)), np.ones((1, 1))))  # integrating model


def _setup_sys_model_id3():
    """SetupMPCtests' `sys` (test/0_test_module.jl:3-5) as LinModel(sys, Ts, i_d=[3]) builds it (src/model/linmodel.jl:165-198):
    zero-order hold for the two manipulated inputs, Tustin for the measured disturbance -- their poles differ, so the minimal
    realisation keeps four states and the default estimator has nx̂ = 6."""
    Ts = 400.0
    a1, b1, c1 = (m[0, 0] for m in zoh_first_order(1.90, 1800.0, Ts))
    a2, b2, c2 = (m[0, 0] for m in zoh_first_order(0.74, 800.0, Ts))

    def tustin(k, tau):
        al = Ts / (2 * tau)
        p = (1 - al) / (1 + al)
        g = k * al / (1 + al)
        return p, g * (1 + p), g  # x+ = p x + d,  y = c x + Dd d  <=>  g (z + 1) / (z - p)
    p1, cd1, dd1 = tustin(1.90, 1800.0)
    p2, cd2, dd2 = tustin(-0.74, 800.0)
    A = np.diag([a1, a2, p1, p2])
    Bu = np.array([[b1, b1], [-b2, b2], [0, 0], [0, 0]])
    Bd = np.array([[0.0], [0.0], [1.0], [1.0]])
    C = np.array([[c1, 0, cd1, 0], [0, c2, 0, cd2]])
    Dd = np.array([[dd1], [dd2]])
    return LinModel(A, Bu, C, Bd=Bd, Dd=Dd, Ts=Ts)


def test_setconstraint_known_answers():
    """test/3_test_predictive_control.jl:259-389 ("LinMPC set constraints"): defaults, every bound and softness keyword in
    its per-sample and whole-horizon form, the softness column of every block of A (construct.jl:999-1199), and the error
    cases (sizes, negative softness, softness / +-Inf pattern frozen after the first moveinput!, softness with Cwt = Inf)."""
    model = _setup_sys_model_id3()
    mpc = LinMPC(model, Hp=1, Hc=1, Wr=np.ones((2, 2)))
    c, nu, ny, nw, nx = mpc.con, 2, 2, 2, 6
    assert mpc.estim.nxhat == nx and mpc.nw == nw
    inf = np.inf

    def blocks(m):
        """softness columns -A[:, end] of (Umin, Umax, DUmin, DUmax, Ymin, Ymax, Wmin, Wmax, xmin, xmax)"""
        sizes = [m.model.nu * m.Hp] * 2 + [m.model.nu * m.Hc] * 2 + [m.model.ny * m.Hp] * 2 + [m.nw * (m.Hp + 1)] * 2 + [m.estim.nxhat] * 2
        out, o = [], 0
        for s in sizes:
            out.append(-m.con.A[o:o + s, -1])
            o += s
        assert o == m.con.A.shape[0]
        return out
    for v, n, val in ((c.U0min, nu, -inf), (c.U0max, nu, inf), (c.DUmin, nu, -inf), (c.DUmax, nu, inf), (c.Y0min, ny, -inf),
                      (c.Y0max, ny, inf), (c.Wmin, 2 * nw, -inf), (c.Wmax, 2 * nw, inf), (c.xhat0min, nx, -inf), (c.xhat0max, nx, inf)):
        assert v.shape == (n,) and (v == val).all()
    b = blocks(mpc)
    for i in (0, 1, 2, 3):
        assert (b[i] == 0.0).all()      # inputs and increments: hard by default
    for i in (4, 5, 6, 7, 8, 9):
        assert (b[i] == 1.0).all()      # outputs, custom rows, terminal states: soft by default
    mpc.setconstraint(umin=[-5, -9.9], umax=[100, 99])
    assert np.allclose(c.U0min, [-5, -9.9]) and np.allclose(c.U0max, [100, 99])
    mpc.setconstraint(dumin=[-5, -10], dumax=[6, 11])
    assert np.allclose(c.DUmin, [-5, -10]) and np.allclose(c.DUmax, [6, 11])
    mpc.setconstraint(ymin=[-6, -11], ymax=[55, 35])
    assert np.allclose(c.Y0min, [-6, -11]) and np.allclose(c.Y0max, [55, 35])
    mpc.setconstraint(wmin=[-7, -12], wmax=[75, 65])
    assert np.allclose(c.Wmin, [-7, -12, -7, -12]) and np.allclose(c.Wmax, [75, 65, 75, 65])
    mpc.setconstraint(xhatmin=[-21, -22, -23, -24, -25, -26], xhatmax=[21, 22, 23, 24, 25, 26])
    assert np.allclose(c.xhat0min, [-21, -22, -23, -24, -25, -26]) and np.allclose(c.xhat0max, [21, 22, 23, 24, 25, 26])
    mpc.setconstraint(c_umin=[0.01, 0.02], c_umax=[0.03, 0.04])
    mpc.setconstraint(c_dumin=[0.05, 0.06], c_dumax=[0.07, 0.08])
    mpc.setconstraint(c_ymin=[1.00, 1.01], c_ymax=[1.02, 1.03])
    mpc.setconstraint(c_wmin=[2.00, 2.01], c_wmax=[2.02, 2.03])
    mpc.setconstraint(c_xhatmin=[0.21, 0.22, 0.23, 0.24, 0.25, 0.26], c_xhatmax=[0.31, 0.32, 0.33, 0.34, 0.35, 0.36])
    expect = [[0.01, 0.02], [0.03, 0.04], [0.05, 0.06], [0.07, 0.08], [1.00, 1.01], [1.02, 1.03],
              [2.00, 2.01, 2.00, 2.01], [2.02, 2.03, 2.02, 2.03],
              [0.21, 0.22, 0.23, 0.24, 0.25, 0.26], [0.31, 0.32, 0.33, 0.34, 0.35, 0.36]]
    for got, want in zip(blocks(mpc), expect):
        assert np.allclose(got, want)

    mpc2 = LinMPC(LinModel(*zoh_first_order(2, 10, 3.0), Ts=3.0), Hp=50, Hc=5, Wr=[[1]])
    c2 = mpc2.con
    r50, r5, r51 = np.arange(1, 51.0), np.arange(1, 6.0), np.arange(1, 52.0)
    mpc2.setconstraint(Umin=-r50 - 1, Umax=r50 + 1)
    assert np.allclose(c2.U0min, -r50 - 1) and np.allclose(c2.U0max, r50 + 1)
    mpc2.setconstraint(DUmin=-r5 - 2, DUmax=r5 + 2)
    assert np.allclose(c2.DUmin, -r5 - 2) and np.allclose(c2.DUmax, r5 + 2)
    mpc2.setconstraint(Ymin=-r50 - 3, Ymax=r50 + 3)
    assert np.allclose(c2.Y0min, -r50 - 3) and np.allclose(c2.Y0max, r50 + 3)
    mpc2.setconstraint(Wmin=-r51 - 4, Wmax=r51 + 4)
    assert np.allclose(c2.Wmin, -r51 - 4) and np.allclose(c2.Wmax, r51 + 4)
    mpc2.setconstraint(C_umin=r50 + 5, C_umax=r50 + 5)
    mpc2.setconstraint(C_dumin=r5 + 6, C_dumax=r5 + 6)
    mpc2.setconstraint(C_ymin=r50 + 7, C_ymax=r50 + 7)
    mpc2.setconstraint(C_wmin=r51 + 8, C_wmax=r51 + 8)
    b2 = blocks(mpc2)
    for i, want in ((0, r50 + 5), (1, r50 + 5), (2, r5 + 6), (3, r5 + 6), (4, r50 + 7), (5, r50 + 7), (6, r51 + 8), (7, r51 + 8)):
        assert np.allclose(b2[i], want)
    mpc2.setconstraint(c_umin=[0], c_umax=[0], c_dumin=[0], c_dumax=[0], c_ymin=[1], c_ymax=[1], c_wmin=[1], c_wmax=[1])

    for kw in ("umin", "umax", "dumin", "dumax", "ymin", "ymax", "wmin", "wmax",
               "c_umin", "c_umax", "c_dumin", "c_dumax", "c_ymin", "c_ymax", "c_wmin", "c_wmax"):
        with pytest.raises(ValueError):       # DimensionMismatch
            mpc.setconstraint(**{kw: [0, 0, 0]})
    for kw in ("c_umin", "c_umax", "c_dumin", "c_dumax", "c_ymin", "c_ymax", "c_wmin", "c_wmax"):
        with pytest.raises(ValueError):       # negative softness
            mpc.setconstraint(**{kw: [-1, -1]})
    mpc.preparestate(model.yop, model.dop)
    mpc.moveinput([0, 0], [0])
    with pytest.raises(RuntimeError):         # softness is frozen after the first solve
        mpc.setconstraint(c_umin=[1, 1], c_umax=[1, 1])
    with pytest.raises(RuntimeError):         # ... and so is the +-Inf pattern
        mpc.setconstraint(umin=[-inf, -inf], umax=[inf, inf])
    mpc3 = LinMPC(model, Cwt=inf)
    for kw in ("c_umin", "c_umax", "c_dumin", "c_dumax", "c_ymin", "c_ymax"):
        with pytest.raises(ValueError):       # ArgumentError: no slack variable
            mpc3.setconstraint(**{kw: [1, 1]})


@pytest.mark.parametrize("kw", [dict(nint_u=[1]), dict(nint_ym=[1])])
def test_steady_kalman_filter_step_disturbance_rejection(kw):
    """test/3_test_predictive_control.jl:177-208 ("LinMPC step disturbance rejection", SteadyKalmanFilter with an input or
    an output integrator): a constant output disturbance of -5 is rejected, the loop settles at ym = r = 15 with u = 2
    (plant gain 5, yop = 10; the controller's own model has no operating point, as in the reference)."""
    plant = LinModel(*zoh_first_order(5, 2, 3.0), Ts=3.0, yop=[10])
    mpc = LinMPC(SteadyKalmanFilter(LinModel(*zoh_first_order(5, 2, 3.0), Ts=3.0), **kw))
    u = ym = None
    for i in range(25):
        ym = plant.evaloutput() - 5
        mpc.preparestate(ym)
        u = mpc.moveinput([15])
        mpc.updatestate(u, ym)
        plant.updatestate(u)
    assert u == pytest.approx([2], abs=1e-2) and ym == pytest.approx([15], abs=1e-2)


def test_linmpc_construction_known_answers():
    """test/3_test_predictive_control.jl:1-91 ("LinMPC construction"): sizes of Ẽ with and without the slack column, the
    weight matrices built from Mwt / Nwt / Lwt / Cwt and given whole (M_Hp, N_Hc, L_Hp), the estimator choices, the
    MultipleShooting decision vector and defect matrix, move-blocking vectors, the custom-constraint block matrices, and
    the constructor's error cases (ArgumentError / DimensionMismatch -> ValueError)."""
    from oracle.linmpc_ms import LinMPCMultipleShooting
    from oracle.mhe import KalmanFilter
    model = _setup_sys_model_id3()
    nu, ny, nd = model.nu, model.ny, model.nd
    mpc1 = LinMPC(model, Hp=15)
    assert isinstance(mpc1.estim, SteadyKalmanFilter) and mpc1.Etilde.shape[0] == 15 * ny
    assert LinMPC(model, Hc=4, Cwt=np.inf).Etilde.shape[1] == 4 * nu
    mpc3 = LinMPC(model, Hc=4, Cwt=1e4)
    assert mpc3.Etilde.shape[1] == 4 * nu + 1 and mpc3.Ntilde_Hc[-1, -1] == 1e4
    assert np.array_equal(LinMPC(model, Mwt=[1, 2], Hp=15).M_Hp, np.diag(np.tile([1.0, 2.0], 15)))
    assert np.array_equal(LinMPC(model, Nwt=[3, 4], Cwt=1e3, Hc=5).Ntilde_Hc, np.diag(np.r_[np.tile([3.0, 4.0], 5), 1e3]))
    assert np.array_equal(LinMPC(model, Lwt=[0, 1], Hp=15).L_Hp, np.diag(np.tile([0.0, 1.0], 15)))
    assert isinstance(LinMPC(KalmanFilter(model)).estim, KalmanFilter)
    mpc9 = LinMPC(model, nint_u=[1, 1], nint_ym=[0, 0])
    assert list(mpc9.estim.nint_u) == [1, 1] and list(mpc9.estim.nint_ym) == [0, 0]
    d20 = np.diag(np.linspace(1.01, 1.2, 20))
    assert np.allclose(LinMPC(model, M_Hp=d20).M_Hp, d20)
    d4 = np.diag([0.1, 0.11, 0.12, 0.13])
    assert np.allclose(LinMPC(model, N_Hc=d4, Cwt=np.inf).Ntilde_Hc, d4)
    l20 = np.diag(np.linspace(0.001, 0.02, 20))
    assert np.allclose(LinMPC(model, L_Hp=l20).L_Hp, l20)
    model2 = LinModel(0.5 * np.ones((1, 1)), np.ones((1, 1)), np.ones((1, 1)), Ts=1.0)
    mpc14 = LinMPCMultipleShooting(model2)
    assert mpc14.Ztilde.size == model2.nu * mpc14.Hc + mpc14.estim.nxhat * mpc14.Hp + mpc14.neps
    assert mpc14.Aeq.shape[0] == mpc14.estim.nxhat * mpc14.Hp
    for Hc in ([1, 2, 3], [1, 2, 3, 6, 6, 6]):  # a block is appended / the blocks past Hp are dropped
        m = LinMPC(model, Hc=Hc, Hp=10, Cwt=np.inf)
        assert m.Hc == 4 and m.Ptilde_u.shape == (10 * nu, 4 * nu)
    rd = lambda W, n: np.kron(np.eye(n), W)
    mpc17 = LinMPC(model, Wy=np.ones((3, ny)))
    n1 = mpc17.Hp + 1
    assert np.array_equal(mpc17.Wbar_y, rd(np.ones((3, ny)), n1))
    assert mpc17.Wbar_u.shape == (3 * n1, nu * n1) and not mpc17.Wbar_u.any()
    assert mpc17.Wbar_d.shape == (3 * n1, nd * n1) and not mpc17.Wbar_d.any()
    assert mpc17.Wbar_r.shape == (3 * n1, ny * n1) and not mpc17.Wbar_r.any()
    Wy, Wu, Wd, Wr = np.ones((2, ny)), 2 * np.ones((2, nu)), 3 * np.ones((2, nd)), 0.5 * np.ones((2, ny))
    mpc18 = LinMPC(model, Wy=Wy, Wu=Wu, Wd=Wd, Wr=Wr)
    for got, W in ((mpc18.Wbar_y, Wy), (mpc18.Wbar_u, Wu), (mpc18.Wbar_d, Wd), (mpc18.Wbar_r, Wr)):
        assert np.array_equal(got, rd(W, mpc18.Hp + 1))
    for kw in (dict(Hp=0), dict(Hc=0), dict(Hp=1, Hc=2), dict(Mwt=[1]), dict(Nwt=[1]), dict(Lwt=[1]), dict(Cwt=[1]),
               dict(Mwt=[-1, 1]), dict(Nwt=[-1, 1]), dict(Lwt=[-1, 1]), dict(Cwt=-1),
               dict(Wy=np.ones((2, ny + 1))), dict(Wu=np.ones((2, nu - 1))), dict(Wd=np.ones((2, nd + 1))),
               dict(Wr=np.ones((2, ny - 1))), dict(Wy=np.ones((2, ny)), Wu=np.ones((3, nu)))):
        with pytest.raises((ValueError, TypeError)):
            LinMPC(model, **kw)


def test_explicitmpc_known_answers():
    """test/3_test_predictive_control.jl:640-781 (ExplicitMPC: "moves and getinfo", "step disturbance rejection",
    "constraints", "set model"): the unconstrained closed form Z̃ = -H̃⁻¹q̃ (explicitmpc.jl:209) -- what stage 2 of the CUDA
    step kernels computes with the cached factor."""
    from oracle.linmpc import InternalModel
    from oracle.mhe import KalmanFilter
    model = first_order(5, 2, 3.0)
    mpc1 = ExplicitMPC(model, Nwt=[0], Hp=1000, Hc=1)
    r, y = [5], [0]
    mpc1.preparestate(y)
    assert mpc1.moveinput(r) == pytest.approx([1], abs=1e-2)
    u = mpc1.moveinput(r, lastu=[-1])
    assert u == pytest.approx([1], abs=1e-2)
    info = mpc1.getinfo()
    assert info["u"] == pytest.approx(u) and info["Yhat"][-1] == pytest.approx(5, abs=1e-2)
    assert info["DU"] == pytest.approx([2.0], abs=1e-2)
    mpc3 = ExplicitMPC(model, Mwt=[0], Nwt=[0], Lwt=[1])
    mpc3.preparestate(y)
    assert mpc3.moveinput([0], Rhat_u=np.full(mpc3.Hp, 12.0)) == pytest.approx([12], abs=1e-2)
    mpc4 = ExplicitMPC(LinModel(0.5 * np.ones((1, 1)), np.ones((1, 1)), np.ones((1, 1)), Ts=1.0))
    mpc4.preparestate(y)
    assert mpc4.moveinput([0]) == pytest.approx([0.0], abs=1e-12)
    mpc5 = ExplicitMPC(model, Hp=10, Hc=[1, 2, 3, 4], Nwt=[10])
    mpc5.preparestate(y)
    mpc5.moveinput(r)
    assert np.diff(mpc5.getinfo()["U"])[[1, 3, 4, 6, 7, 8]] == pytest.approx(np.zeros(6), abs=1e-9)
    # step disturbance rejection (:676-727): InternalModel, input integrator, output integrator
    for make in (lambda: InternalModel(first_order(5, 2, 3.0, yop=[10])),
                 lambda: SteadyKalmanFilter(first_order(5, 2, 3.0), nint_u=[1]),
                 lambda: SteadyKalmanFilter(first_order(5, 2, 3.0), nint_ym=[1])):
        plant, mpc = first_order(5, 2, 3.0, yop=[10]), ExplicitMPC(make())
        u = ym = None
        for i in range(25):
            ym = plant.evaloutput() - 5
            mpc.preparestate(ym)
            u = mpc.moveinput([15])
            mpc.updatestate(u, ym)
            plant.updatestate(u)
        assert u == pytest.approx([2], abs=1e-2) and ym == pytest.approx([15], abs=1e-2)
    with pytest.raises(RuntimeError):  # :743-748
        ExplicitMPC(_setup_sys_model_id3(), Hp=1, Hc=1).setconstraint(umin=[0.0, 0.0])
    # set model (:750-781)
    mpc = ExplicitMPC(KalmanFilter(first_order(5, 2, 3.0, yop=[10], uop=[1])), Nwt=[0], Hp=1000, Hc=1)
    assert np.array_equal(mpc.Yop, np.full(1000, 10.0)) and np.array_equal(mpc.Uop, np.full(1000, 1.0))
    mpc.preparestate([10])
    assert mpc.moveinput([15]) == pytest.approx([2], abs=1e-2)
    assert mpc.lastu0 == pytest.approx([2 - 1], abs=1e-2)
    mpc.setmodel(first_order(5, 2, 3.0, yop=[20], uop=[11]))
    assert np.array_equal(mpc.Yop, np.full(1000, 20.0)) and np.array_equal(mpc.Uop, np.full(1000, 11.0))
    assert mpc.lastu0 == pytest.approx([2 - 11], abs=1e-2)
    assert mpc.moveinput([40]) == pytest.approx([15], abs=1e-2)
    mpc.setmodel(first_order(10, 2, 3.0, yop=[20], uop=[11]))
    assert mpc.moveinput([40]) == pytest.approx([13], abs=1e-2)
    mpc.setmodel(Mwt=[100], Nwt=[200], Lwt=[300])
    assert np.array_equal(mpc.M_Hp, np.diag(np.full(1000, 100.0))) and np.array_equal(mpc.Ntilde_Hc, [[200.0]])
    assert np.array_equal(mpc.L_Hp, np.diag(np.full(1000, 300.0)))
    mpc.setmodel(M_Hp=np.diag(np.arange(1, 1001.0)), Ntilde_Hc=[0.1], L_Hp=np.diag(np.arange(1.1, 1000.2)))
    assert np.allclose(mpc.M_Hp, np.diag(np.arange(1, 1001.0))) and np.allclose(mpc.Ntilde_Hc, [[0.1]])
    assert np.allclose(mpc.L_Hp, np.diag(np.arange(1.1, 1000.2)))


def test_steady_kalman_filter_estimator_methods():
    """test/2_test_state_estim.jl:64-127 ("SKF estimator methods"): the observer that feeds x̂0 to moveinput! (and that the
    step kernels can run fused): zero estimate at the operating point, initstate!, setstate!, convergence of the estimated
    output under an input offset and an output offset in the current (direct) and prediction forms, NaN measurements
    skipped (kalman.jl:248-251)."""
    Ts = 400.0
    a1, b1, g1 = (m[0, 0] for m in zoh_first_order(1.90, 1800.0, Ts))
    a2, b2, g2 = (m[0, 0] for m in zoh_first_order(0.74, 800.0, Ts))
    mk = lambda: LinModel(np.diag([a1, a2]), np.array([[b1, b1], [-b2, b2]]), np.diag([g1, g2]), Ts=Ts, uop=[10, 50], yop=[50, 30])
    kf1 = SteadyKalmanFilter(mk(), nint_ym=[1, 1])
    u, y = [10, 50], [50, 30]
    kf1.preparestate(y)
    assert kf1.updatestate(u, y) == pytest.approx(np.zeros(4), abs=1e-12)
    kf1.preparestate(y)
    assert kf1.evaloutput() == pytest.approx([50, 30])
    assert kf1.initstate([10, 50], [50, 30 + 1]) == pytest.approx([0, 0, 0, 1], abs=1e-9)
    # an integrating plant and a first-order one, input integrators, prediction form (:81-85)
    ad, bd, cd = (m[0, 0] for m in zoh_first_order(2, 10, 1.0))
    m2 = LinModel(np.diag([1.0, ad]), np.diag([1.0, bd]), np.diag([1.0, cd]), Ts=1.0)
    kf2 = SteadyKalmanFilter(m2, nint_u=[1, 1], direct=False)
    x = kf2.initstate([10, 3], [0.5, 6 + 0.1])
    assert kf2.evaloutput() == pytest.approx([0.5, 6.1])
    assert kf2.updatestate([10, 3], [0.5, 6 + 0.1]) == pytest.approx(x, abs=1e-9)
    kf1.setstate([1, 2, 3, 4])
    assert kf1.xhat0 == pytest.approx([1, 2, 3, 4])
    for est, prep in ((kf1, True), (SteadyKalmanFilter(mk(), nint_u=[1, 1], direct=False), False)):
        for uu, ym in (([11, 52], [50, 30]), ([10, 50], [51, 32])):
            for _ in range(40):
                est.preparestate(ym)
                est.updatestate(uu, ym)
            if prep:
                est.preparestate(ym)
            assert est.evaloutput() == pytest.approx(ym, abs=1e-3)
    kf3 = SteadyKalmanFilter(LinModel(0.5 * np.ones((1, 1)), np.ones((1, 1)), np.ones((1, 1)), Ts=1.0))
    kf3.preparestate([0])
    assert kf3.updatestate([0], [0]) == pytest.approx([0, 0], abs=1e-12)
    kf4 = SteadyKalmanFilter(mk(), nint_ym=[1, 1], direct=True)
    kf4.xhat0[:] = 7
    kf4.preparestate([55, np.nan])
    assert (kf4.xhat0 == 7).all()
    kf5 = SteadyKalmanFilter(mk(), nint_ym=[1, 1], direct=False)
    kf5.updatestate([10, 50], [55, np.nan])
    assert np.isfinite(kf5.xhat0).all()


def test_linmpc_other_methods_and_moveinput_argument_sizes():
    """test/3_test_predictive_control.jl:239-257 ("LinMPC other methods": initstate!, setstate!(mpc, x̂, P̂) and a period at
    the operating point through the controller's pass-throughs, KalmanFilter estimator) and :153-157 (moveinput!'s
    validate_args, construct.jl:702-710: DimensionMismatch -> ValueError)."""
    from oracle.mhe import KalmanFilter
    Ts = 400.0
    a1, b1, g1 = (m[0, 0] for m in zoh_first_order(1.90, 1800.0, Ts))
    a2, b2, g2 = (m[0, 0] for m in zoh_first_order(0.74, 800.0, Ts))
    model = LinModel(np.diag([a1, a2]), np.array([[b1, b1], [-b2, b2]]), np.diag([g1, g2]), Ts=Ts, uop=[10, 50], yop=[50, 30])
    mpc1 = LinMPC(KalmanFilter(model))
    assert mpc1.initstate([10, 50], [50, 30 + 1]) == pytest.approx([0, 0, 0, 1], abs=1e-9)
    mpc1.setstate([1, 2, 3, 4], np.diag([0.1, 0.2, 0.3, 0.4]))
    assert mpc1.estim.xhat0 == pytest.approx([1, 2, 3, 4]) and np.allclose(mpc1.estim.Phat, np.diag([0.1, 0.2, 0.3, 0.4]))
    mpc1.setstate([0, 0, 0, 0], np.diag(np.r_[np.full(2, 0.5), 1, 1]) ** 2)
    mpc1.preparestate([50, 30])
    mpc1.updatestate(model.uop, [50, 30])
    assert mpc1.estim.xhat0 == pytest.approx(np.zeros(4), abs=1e-12)
    with pytest.raises(TypeError):  # updatestate!(mpc1, [0, 0]) without ym: ArgumentError in the reference
        mpc1.updatestate([0, 0])
    lin = LinModel(*zoh_first_order(5, 2, 3.0), Ts=3.0, yop=[10])
    m = LinMPC(lin, Nwt=[0], Hp=1000, Hc=1)
    m.preparestate([10])
    for kw in (dict(ry=[0, 0, 0]), dict(ry=[0], d=[0, 0]), dict(Dhat=np.zeros(m.Hp + 1)), dict(Rhat_y=np.zeros(m.Hp + 1)),
               dict(Rhat_u=np.zeros(m.Hp + 1))):
        with pytest.raises(ValueError):
            m.moveinput(**kw)


def test_internalmodel_construction_and_setmodel():
    """test/2_test_state_estim.jl:406-475 ("IM construction": sizes for the default integrators and a measured-output
    subset, a user stochastic model kept as given, the error cases) and :523-547 ("IM set model")."""
    from oracle.linmpc import InternalModel
    Ts = 400.0
    a1, b1, g1 = (m[0, 0] for m in zoh_first_order(1.90, 1800.0, Ts))
    a2, b2, g2 = (m[0, 0] for m in zoh_first_order(0.74, 800.0, Ts))
    linmodel = LinModel(np.diag([a1, a2]), np.array([[b1, b1], [-b2, b2]]), np.diag([g1, g2]), Ts=Ts)
    im1 = InternalModel(linmodel)
    assert (len(im1.i_ym), linmodel.ny - len(im1.i_ym), im1.nxs, im1.nxhat) == (2, 0, 2, 2)
    linmodel2 = _setup_sys_model_id3()
    im2 = InternalModel(linmodel2, i_ym=[1])                        # the reference's i_ym=[2] (1-based)
    assert (len(im2.i_ym), linmodel2.ny - len(im2.i_ym), im2.nxs, im2.nxhat) == (1, 1, 1, 4)
    rng = np.random.default_rng(5)
    As, Bs, Cs, Ds = 0.4 * rng.standard_normal((4, 4)), rng.standard_normal((4, 2)), rng.standard_normal((2, 4)), np.eye(2)
    im5 = InternalModel(linmodel2, stoch_ym=(As, Bs, Cs, Ds))
    assert (im5.nxs, im5.nxhat) == (4, 4)
    assert np.array_equal(im5.As, As) and np.array_equal(im5.Bs, Bs) and np.array_equal(im5.Cs, Cs) and np.array_equal(im5.Ds, Ds)
    with pytest.raises(ValueError):   # integrating / unstable plant model
        InternalModel(LinModel(np.diag([0.5, -0.5, 1.5]), np.ones((3, 1)), np.eye(3), Ts=1.0))
    with pytest.raises(ValueError):   # one stochastic output for two measured outputs
        InternalModel(linmodel, stoch_ym=([[1.0]], [[1.0]], [[1.0]], [[1.0]]))
    with pytest.raises(ValueError):   # no direct transmission
        InternalModel(linmodel, stoch_ym=(np.eye(2), np.eye(2), np.eye(2), np.zeros((2, 2))))
    # set model
    lin = lambda a, uop, yop, xop: LinModel([[a]], [[0.3]], [[1.0]], Ts=10.0, uop=[uop], yop=[yop], xop=[xop], fop=[xop])
    im = InternalModel(lin(0.5, 2.0, 50.0, 3.0))
    assert np.allclose(im.Ahat, [[0.5]])
    im.preparestate([50.0])
    assert im.evaloutput() == pytest.approx([50.0])
    im.preparestate([50.0])
    assert im.updatestate([2.0], [50.0]) == pytest.approx([3.0])
    im.setmodel(lin(0.2, 3.0, 55.0, 3.0))
    assert np.allclose(im.Ahat, [[0.2]])
    im.preparestate([55.0])
    assert im.evaloutput() == pytest.approx([55.0])
    im.preparestate([55.0])
    assert im.updatestate([3.0], [55.0]) == pytest.approx([3.0])
    im.setmodel(lin(0.2, 3.0, 55.0, 8.0))
    assert im.xhat0 == pytest.approx([3.0 - 8.0])


def test_manual_estimator_construction_and_methods():
    """test/2_test_state_estim.jl:1889-1960 ("Manual construction", "Manual estimator methods", LinModel parts): the
    ManualEstimator is the reference's own seam for an externally supplied x̂0 (src/estimator/manual.jl:60-64,150-154) -- the
    one the C ABI's `xhat0` argument stands for: same augmentation as the other estimators, preparestate! / updatestate!
    leave the state alone, setstate! sets it."""
    Ts = 400.0
    a1, b1, g1 = (m[0, 0] for m in zoh_first_order(1.90, 1800.0, Ts))
    a2, b2, g2 = (m[0, 0] for m in zoh_first_order(0.74, 800.0, Ts))
    linmodel = LinModel(np.diag([a1, a2]), np.array([[b1, b1], [-b2, b2]]), np.diag([g1, g2]), Ts=Ts)
    m1 = ManualEstimator(linmodel)
    assert (len(m1.i_ym), m1.nxs, m1.nxhat, list(m1.nint_ym)) == (2, 2, 4, [1, 1])
    m2 = ManualEstimator(_setup_sys_model_id3(), i_ym=[1])           # the reference's i_ym=[2] (1-based)
    assert (len(m2.i_ym), m2.nxs, m2.nxhat, list(m2.nint_ym)) == (1, 1, 5, [1])
    m3 = ManualEstimator(linmodel, nint_ym=0)
    assert (m3.nxs, m3.nxhat, list(m3.nint_ym)) == (0, 2, [0, 0])
    m4 = ManualEstimator(linmodel, nint_ym=[2, 2])
    assert (m4.nxs, m4.nxhat) == (4, 6)
    m5 = ManualEstimator(linmodel, nint_u=[1, 1])
    assert (m5.nxs, m5.nxhat, list(m5.nint_u), list(m5.nint_ym)) == (2, 4, [1, 1], [0, 0])
    m1.preparestate([50, 30])
    assert (m1.xhat0 == 0).all()
    m1.updatestate([11, 52], [50, 30])
    assert (m1.xhat0 == 0).all()
    m1.setstate([1, 2, 3, 4])
    assert m1.xhat0 == pytest.approx([1, 2, 3, 4])


def test_steady_kalman_filter_construction():
    """test/2_test_state_estim.jl:1-62 ("SKF construction"): augmentation sizes for the default and the given integrators
    (default_nint leaves an integrating output without one, estimator/construct.jl), covariances from standard deviations,
    and the error cases (sizes, negative counts, unobservable augmentations)."""
    Ts = 400.0
    a1, b1, g1 = (m[0, 0] for m in zoh_first_order(1.90, 1800.0, Ts))
    a2, b2, g2 = (m[0, 0] for m in zoh_first_order(0.74, 800.0, Ts))
    linmodel = LinModel(np.diag([a1, a2]), np.array([[b1, b1], [-b2, b2]]), np.diag([g1, g2]), Ts=Ts)
    k1 = SteadyKalmanFilter(linmodel)
    assert (len(k1.i_ym), k1.nxs, k1.nxhat, list(k1.nint_ym)) == (2, 2, 4, [1, 1])
    linmodel2 = _setup_sys_model_id3()
    k2 = SteadyKalmanFilter(linmodel2, i_ym=[1])
    assert (len(k2.i_ym), k2.nxs, k2.nxhat, list(k2.nint_ym)) == (1, 1, 5, [1])
    k3 = SteadyKalmanFilter(linmodel, nint_ym=0)
    assert (k3.nxs, k3.nxhat, list(k3.nint_ym)) == (0, 2, [0, 0])
    k4 = SteadyKalmanFilter(linmodel, nint_ym=[2, 2])
    assert (k4.nxs, k4.nxhat) == (4, 6)
    k5 = SteadyKalmanFilter(linmodel2, sigmaQ=[1, 2, 3, 4], sigmaQint_ym=[5, 6], sigmaR=[7, 8])
    assert np.array_equal(k5.Qhat, np.diag([1.0, 4, 9, 16, 25, 36])) and np.array_equal(k5.Rhat, np.diag([49.0, 64]))
    # append(1/s, 1/(10s+1), 1/(-s+1)) at Ts = 0.1 (zero-order hold): an integrator, a stable and an unstable pole
    ast, bst, cst = (m[0, 0] for m in zoh_first_order(1, 10, 0.1))
    au = np.exp(0.1)                                     # 1/(1 - s) = -1/(s - 1): x' = x + u, y = -x
    linmodel3 = LinModel(np.diag([1.0, ast, au]), np.diag([0.1, bst, au - 1.0]), np.diag([1.0, cst, -1.0]), Ts=0.1)
    k6 = SteadyKalmanFilter(linmodel3)
    assert (k6.nxs, k6.nxhat, list(k6.nint_ym)) == (2, 5, [0, 1, 1])
    k7 = SteadyKalmanFilter(linmodel, nint_u=[1, 1])
    assert (k7.nxs, k7.nxhat, list(k7.nint_u), list(k7.nint_ym)) == (2, 4, [1, 1], [0, 0])
    for kw in (dict(nint_ym=[1, 1, 1]), dict(nint_ym=[-1, 0]), dict(nint_ym=0, sigmaQ=[1]), dict(nint_ym=0, sigmaR=[1, 1, 1]),
               dict(nint_u=[1, 1], nint_ym=[1, 1])):
        with pytest.raises(ValueError):
            SteadyKalmanFilter(linmodel, **kw)
    with pytest.raises(ValueError):                      # an integrator on the already integrating output: unobservable
        SteadyKalmanFilter(linmodel3, nint_ym=[1, 0, 0])
    with pytest.raises(ValueError):                      # unobservable plant mode
        SteadyKalmanFilter(LinModel([[1, 0], [0, 1.5]], [[1], [0]], [[1, 0]], Ts=1.0), nint_ym=[1])
