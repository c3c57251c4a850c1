"""GPU parity tests: the CUDA step (through the C ABI) against the oracle on the same inputs.

Tolerances (stated, fp64):
  * assembly outputs F, q̃, r (initpred!):            1e-10 relative
  * steps without active constraints (u, Z̃, J):       1e-9
  * steps with active constraints: Z̃ within 5e-6*(1+|Z̃|), u within 5e-6, J within 1e-8 relative (measured worst: 1.4e-6 / 7e-9)
    (the interior-point iterate stops at relative KKT residual 1e-11 / duality gap 1e-14, or at the fp64 floor; the error
    of Z̃ is proportional to the gap times the flatness of the objective.  The reference's own
    default solver OSQP stops at 1e-3 and its tests assert 1e-2..1e-1 -- SURVEY section 0 item 5).
"""
import numpy as np
import pytest

from oracle import qp
from oracle.linmpc import LinModel, LinMPC, zoh_first_order
from helpers import batch_from_oracle, c1_controllers, push_constraints

pytestmark = pytest.mark.gpu

TOL_Z, TOL_J = 5e-6, 1e-8


def closed_loop_compare(mpcs, plants, b, rng, steps, switch=25, check_info=True):
    N = len(mpcs)
    ny = mpcs[0].model.ny
    r = rng.choice([-1.0, 1.0], (N, ny))
    worst = dict(z=0.0, u=0.0, J=0.0, F=0.0, q=0.0)
    n_active = 0
    for k in range(steps):
        if k % switch == 0:
            r = rng.choice([-1.0, 1.0], (N, ny))
        ys = [p.evaloutput() for p in plants]
        for m, y in zip(mpcs, ys):
            m.preparestate(y)
        xhat0 = np.stack([m.estim.xhat0 for m in mpcs])
        b.lastu0[:] = np.stack([m.lastu0 for m in mpcs])
        u_gpu = b.step(xhat0, ry=r).copy()
        for i, m in enumerate(mpcs):
            u = m.moveinput(r[i])
            assert m.last_status == qp.OPTIMAL
            assert b.status[i] == 0, (k, i, b.status[i], b.iters[i])
            tz = TOL_Z if b.iters[i] > 0 else 1e-9
            n_active += b.iters[i] > 0
            ez = np.abs(b.Ztilde[i] - m.Ztilde).max() / (1 + np.abs(m.Ztilde).max())
            assert ez < tz, (k, i, ez, b.iters[i])
            assert np.abs(u_gpu[i] - u).max() < tz * (1 + np.abs(u).max())
            Jo = m.getinfo()["J"]
            ej = abs(b.J[i] - Jo) / (1 + abs(Jo))
            assert ej < TOL_J, (k, i, ej)
            worst["z"], worst["J"] = max(worst["z"], ez), max(worst["J"], ej)
            m.updatestate(u, ys[i])
            plants[i].updatestate(u)
        if check_info and k % 7 == 0:
            info = b.getinfo()
            for i, m in enumerate(mpcs):
                assert np.abs(info["F"][i] - m.F).max() < 1e-10 * (1 + np.abs(m.F).max())
                assert np.abs(info["qtilde"][i] - m.qtilde).max() < 1e-10 * (1 + np.abs(m.qtilde).max())
                assert abs(info["r"][i] - m.r) < 1e-10 * (1 + abs(m.r))
                oi = m.getinfo()
                tz = TOL_Z if b.iters[i] > 0 else 1e-9
                assert np.abs(info["Yhat0"][i] + m.Yop - oi["Yhat"]).max() < 10 * tz * (1 + np.abs(oi["Yhat"]).max())
                assert np.abs(info["U0"][i] + m.Uop - oi["U"]).max() < 10 * tz * (1 + np.abs(oi["U"]).max())
    return worst, n_active


@pytest.mark.parametrize("team", [0, 8, 32, 128])
def test_c1_closed_loop_parity(team):
    """Config C1 recipe (2x2 plants, Hp=20, Hc=5, hard u box + soft ymax) at a small batch."""
    mpcs, plants, rng = c1_controllers(24, seed=1)
    b = batch_from_oracle(mpcs, team=team)
    worst, n_active = closed_loop_compare(mpcs, plants, b, rng, steps=40)
    assert n_active > 100  # the constraints really were active
    print("C1 worst", worst, "active solves", n_active, b.launch_info())


@pytest.mark.parametrize("nu,ny,Hc,nt", [(1, 1, 2, 3), (2, 2, 2, 5), (2, 2, 3, 7), (2, 2, 4, 9), (3, 2, 4, 13), (3, 3, 5, 16), (2, 2, 7, 16)])
def test_warp_kernel_every_specialisation(nu, ny, Hc, nt):
    """Every compiled size of the warp-per-controller kernel (NT = 3 ... 16; C1 above is NT = 11), odd and even: the
    LDL' factorisation's last-column case and the two-variables-per-step substitutions differ between them.  The decision
    vector has nu*Hc + 1 entries (NT is the smallest specialisation that holds it: 15 rides on 16)."""
    mpcs, plants, rng = c1_controllers(8, seed=10 + nt + Hc, nx=4, nu=nu, ny=ny, Hp=12, Hc=Hc)
    b = batch_from_oracle(mpcs)
    assert nt - 3 < nu * Hc + 1 <= nt
    worst, n_active = closed_loop_compare(mpcs, plants, b, rng, steps=20, switch=10)
    assert n_active > 10
    info = b.launch_info()  # (the launch is configured by the first step)
    assert info["team"] == 32 and info["teams_per_cta"] == 1 and info["pd_in_smem"] > 0  # really the warp kernel
    print("NT", nt, "worst", worst, "active solves", n_active, info)


def test_c2_shapes_parity():
    """Config C2 recipe (4x4, nx=8, Hp=30, Hc=10 -> n=41)."""
    mpcs, plants, rng = c1_controllers(6, seed=2, nx=8, nu=4, ny=4, Hp=30, Hc=10)
    b = batch_from_oracle(mpcs)
    worst, n_active = closed_loop_compare(mpcs, plants, b, rng, steps=12, switch=6)
    assert n_active > 10
    info = b.launch_info()
    assert info["team"] == 64  # n <= 48: two warps per controller, packed Hessian left in L2
    print("C2 worst", worst, info)


@pytest.mark.parametrize("shape", ["n81", "c4"])
def test_mid_and_large_shapes_parity(shape):
    """CTA-team paths of the general kernel: n = 81 (128 threads, blocked DMMA Cholesky, 16x16-block DMMA Hessian build)
    and the config C4 recipe (8x8 plant, Hp=50, Hc=20 -> n = 161, 256 threads, packed Hessian left in HBM/L2, 32x32-block
    DMMA build; hard u box + hard du box + soft ymin/ymax), warm-started from period to period."""
    from helpers import random_plant
    rng = np.random.default_rng(44)
    nx, nu, ny, Hp, Hc = (8, 8, 4, 20, 10) if shape == "n81" else (16, 8, 8, 50, 20)
    mpcs, plants = [], []
    for _ in range(3):
        m = random_plant(rng, nx, nu, ny)
        mpc = LinMPC(m, Hp=Hp, Hc=Hc, Cwt=1e5)
        if shape == "c4":
            mpc.setconstraint(umin=[-1.0] * nu, umax=[1.0] * nu, dumin=[-0.2] * nu, dumax=[0.2] * nu,
                              ymin=[-1.2] * ny, ymax=[0.8] * ny)
        else:
            mpc.setconstraint(umin=[-1.0] * nu, umax=[1.0] * nu, ymax=[0.8] * ny)
        mpcs.append(mpc)
        plants.append(LinModel(m.A, m.Bu, m.C))
    b = batch_from_oracle(mpcs)
    worst, n_active = closed_loop_compare(mpcs, plants, b, rng, steps=6, switch=3)
    assert n_active > 6
    info = b.launch_info()
    assert info["team"] == (128 if shape == "n81" else 256)
    print(shape, "worst", worst, info)


def test_reference_known_answers_on_gpu():
    """test/3_test_predictive_control.jl:93-106 through the CUDA path (Hp=1000, Hc=1, Nwt=0)."""
    A, B, C = zoh_first_order(5, 2, 3.0)
    linmodel = LinModel(A, B, C, Ts=3.0, yop=[10])
    for Cwt in (1e5, np.inf):
        mpc = LinMPC(linmodel, Nwt=[0], Hp=1000, Hc=1, Cwt=Cwt)
        b = batch_from_oracle([mpc], shared_model=True)
        mpc.preparestate([10])
        u = b.step(mpc.estim.xhat0[None], ry=np.array([[15.0]]))
        assert u[0] == pytest.approx([1], abs=1e-2)
        b.lastu0[:] = -1.0 - 0.0  # lastu = -1 (uop = 0)
        u = b.step(mpc.estim.xhat0[None], ry=np.array([[15.0]]))
        assert u[0] == pytest.approx([1], abs=1e-2)
        info = b.getinfo()
        assert info["DU"][0] == pytest.approx([2.0], abs=1e-2)
        assert info["Yhat0"][0][-1] + 10 == pytest.approx(15, abs=1e-2)


@pytest.mark.parametrize("Cwt", [1e5, np.inf])
def test_constraint_violation_sequence(Cwt):
    """test/3_test_predictive_control.jl:391-464 (soft and hard), GPU vs oracle at every call,
    with bounds updated between calls by setconstraint! (values change, +-Inf pattern frozen)."""
    A, B, C = zoh_first_order(2, 10, 3.0)
    mpc = LinMPC(LinModel(A, B, C, Ts=3.0), Hp=50, Hc=5, Cwt=Cwt)
    mpc.setconstraint(xhatmin=[-1e6, -np.inf], xhatmax=[1e6, np.inf], umin=[-10], umax=[10],
                      dumin=[-15], dumax=[15], ymin=[-100], ymax=[100])
    if np.isfinite(Cwt):
        mpc.setconstraint(c_xhatmin=[1, 1], c_xhatmax=[1, 1], c_umin=[0.1], c_umax=[0.1],
                          c_dumin=[0.1], c_dumax=[0.1], c_ymin=[1], c_ymax=[1])
    b = batch_from_oracle([mpc], shared_model=True)
    mpc.preparestate([0])

    def both(r, key, expect, tol=1e-1):
        push_constraints(b, [mpc])
        b.lastu0[:] = mpc.lastu0
        b.Ztilde[:] = mpc.Ztilde
        b.step(mpc.estim.xhat0[None], ry=np.array([r], dtype=float))
        mpc.moveinput(r)
        assert b.status[0] == 0 and mpc.last_status == qp.OPTIMAL
        assert np.abs(b.Ztilde[0] - mpc.Ztilde).max() < TOL_Z * (1 + np.abs(mpc.Ztilde).max()), (r, key)
        gi, oi = b.getinfo(), mpc.getinfo()
        val = dict(U=gi["U0"][0], DU=gi["DU"][0], Yhat=gi["Yhat0"][0], xhatend=gi["xhat0end"][0])[key]
        assert abs(gi["J"][0] - oi["J"]) < TOL_J * (1 + abs(oi["J"]))
        return val

    mpc.setconstraint(umin=[-3], umax=[4])
    assert np.allclose(both([-100], "U", -3), -3, atol=1e-1)
    assert np.allclose(both([100], "U", 4), 4, atol=1e-1)
    mpc.setconstraint(umin=[-10], umax=[10], dumin=[-1.5], dumax=[1.25])
    assert np.allclose(both([-100], "DU", -1.5), -1.5, atol=1e-1)
    assert np.allclose(both([100], "DU", 1.25), 1.25, atol=1e-1)
    mpc.setconstraint(dumin=[-15], dumax=[15], ymin=[-0.5], ymax=[0.9])
    assert np.allclose(both([-100], "Yhat", -0.5), -0.5, atol=1e-1)
    assert np.allclose(both([100], "Yhat", 0.9), 0.9, atol=1e-1)
    mpc.setconstraint(ymin=[-100], ymax=[100])
    mpc.setconstraint(Ymin=np.r_[-0.5, np.full(49, -100.0)], Ymax=np.r_[0.9, np.full(49, 100.0)])
    y = both([-10], "Yhat", None)
    assert y[0] == pytest.approx(-0.5, abs=1e-1) and y[-1] == pytest.approx(-10, abs=1e-1)
    y = both([10], "Yhat", None)
    assert y[0] == pytest.approx(0.9, abs=1e-1) and y[-1] == pytest.approx(10, abs=1e-1)
    mpc.setconstraint(ymin=[-100], ymax=[100], xhatmin=[-1e-6, -np.inf], xhatmax=[1e-6, np.inf])
    assert both([-100], "xhatend", 0)[0] == pytest.approx(0, abs=1e-1)
    assert both([100], "xhatend", 0)[0] == pytest.approx(0, abs=1e-1)
    # construct.jl:548-551: the +-Inf pattern is frozen after the first step
    import mpc_b200
    with pytest.raises(mpc_b200.BmpcError) as ei:
        b.set_constraints(U0max=np.full((1, 50), 10.0))
    assert ei.value.code == -3


def test_infeasible_returns_shifted_solution():
    """test/3_test_predictive_control.jl:142-150: umin > umax -> error status -> Z̃ = shifted last."""
    A, B, C = zoh_first_order(5, 2000, 3000.0)
    mpc = LinMPC(LinModel(A, B, C, Ts=3000.0), Hp=4, Hc=3, Cwt=np.inf).setconstraint(umin=[+1], umax=[-1])
    b = batch_from_oracle([mpc], shared_model=True)
    b.Ztilde[:] = [[0.3, 0.2, 0.1]]
    b.lastu0[:] = 0.5
    u = b.step(np.zeros((1, mpc.estim.nxhat)), ry=np.zeros((1, 1)))
    assert b.status[0] == 2
    assert b.Ztilde[0] == pytest.approx([0.2, 0.1, 0.0])
    assert u[0] == pytest.approx([0.7])


def test_dense_M_and_L_weights_lqr():
    """LQR equivalence case (test/3_test_predictive_control.jl:498-527): dense M_Hp, Lwt, nint_ym=0."""
    import scipy.linalg
    A = np.array([[0.5, -0.4], [0.6, 0.5]])
    Bm = Cm = np.eye(2)
    P = scipy.linalg.solve_discrete_are(A, Bm, np.eye(2), 0.5 * np.eye(2))
    M_Hp = np.block([[np.eye(4), np.zeros((4, 2))], [np.zeros((2, 4)), P]])
    mpc = LinMPC(LinModel(A, Bm, Cm), Hp=3, Hc=3, M_Hp=M_Hp, Nwt=[0, 0], Lwt=[0.5, 0.5], nint_ym=0)
    plant = LinModel(A, Bm, Cm)
    b = batch_from_oracle([mpc], shared_model=True)
    mpc.setstate([1, 1])
    plant.setstate([1, 1])
    for i in range(10):
        y = plant.evaloutput()
        mpc.preparestate(y)
        b.lastu0[:] = mpc.lastu0
        ug = b.step(mpc.estim.xhat0[None], ry=np.zeros((1, 2))).copy()
        u = mpc.moveinput([0, 0])
        assert ug[0] == pytest.approx(u, abs=1e-10)
        assert b.J[0] == pytest.approx(mpc.getinfo()["J"], rel=1e-10, abs=1e-12)
        mpc.updatestate(u, y)
        plant.updatestate(u)


@pytest.mark.parametrize("team", [0, 128])
def test_all_features_vs_oracle(team):
    """Everything the linear path offers at once, GPU (route B) against the oracle per call: operating points on
    u/y/d/x, measured disturbance with a preview D̂, input setpoints R̂u with Lwt, an output setpoint trajectory R̂y,
    custom move blocking, hard u box, soft du box, two-sided soft y bounds and terminal state bounds."""
    from helpers import random_plant
    rng = np.random.default_rng(123)
    mpcs = []
    for i in range(6):
        p = random_plant(rng, nx=3, nu=2, ny=2)
        m = LinModel(p.A, p.Bu, p.C, Bd=rng.standard_normal((3, 1)), Dd=rng.standard_normal((2, 1)) * 0.1,
                     uop=[0.5, -0.2], yop=[1.0, 2.0], dop=[0.3], xop=[0.1, 0.0, -0.1], fop=[0.05, 0.1, 0.0])
        mpc = LinMPC(m, Hp=12, Hc=[1, 2, 3], Lwt=[0.3, 0.1], Cwt=1e4)
        mpc.setconstraint(umin=[-2, -2], umax=[2, 2], dumin=[-0.7, -0.7], dumax=[0.7, 0.7], ymin=[0, 1],
                          ymax=[2.5, 3.5], xhatmin=[-5] * mpc.estim.nxhat, xhatmax=[5] * mpc.estim.nxhat,
                          c_dumin=[0.1, 0.1], c_dumax=[0.1, 0.1])
        mpcs.append(mpc)
    b = batch_from_oracle(mpcs, team=team, with_terminal=True)
    N, Hp = len(mpcs), 12
    worst, n_active = 0.0, 0
    for k in range(8):
        xh = rng.standard_normal((N, mpcs[0].estim.nxhat)) * 0.4
        ry = rng.standard_normal((N, 2)) * 0.5 + np.array([1.0, 2.0])
        Ry = np.tile(ry, (1, Hp)) + 0.1 * rng.standard_normal((N, 2 * Hp)) if k % 2 else None
        Ru = np.tile(np.array([0.5, -0.2]), (N, Hp)) + 0.2 * rng.standard_normal((N, 2 * Hp))
        d = 0.3 + 0.2 * rng.standard_normal((N, 1))
        Dh = np.tile(d, (1, Hp)) + 0.05 * rng.standard_normal((N, Hp))
        for i, m in enumerate(mpcs):
            m.estim.xhat0 = xh[i].copy()
        b.lastu0[:] = np.stack([m.lastu0 for m in mpcs])
        b.step(xh, ry=ry, Rhat_y=Ry, Rhat_u=Ru,
               d0=d - 0.3, Dhat0=None if k == 6 else Dh - 0.3)  # k = 6: default D̂ = repeat(d, Hp)
        for i, m in enumerate(mpcs):
            u = m.moveinput(ry[i], d=d[i], Dhat=None if k == 6 else Dh[i], Rhat_y=None if Ry is None else Ry[i], Rhat_u=Ru[i])
            assert b.status[i] == 0 and m.last_status == qp.OPTIMAL, (k, i, b.status[i], b.iters[i])
            ez = np.abs(b.Ztilde[i] - m.Ztilde).max() / (1 + np.abs(m.Ztilde).max())
            tz = TOL_Z if b.iters[i] > 0 else 1e-9
            assert ez < tz, (k, i, ez, b.iters[i])
            assert np.abs(b.u[i] - u).max() < tz * (1 + np.abs(u).max())
            n_active += b.iters[i] > 0
            worst = max(worst, ez)
        if k % 3 == 0:  # getinfo with nd > 0: x̂end carries gx̂ d0 + jx̂ D̂0, Ŷ carries G d0 + J D̂0
            gi = b.getinfo()
            for i, m in enumerate(mpcs):
                oi = m.getinfo()
                tz = 10 * (TOL_Z if b.iters[i] > 0 else 1e-9)
                xe = gi["xhat0end"][i] + m.estim.xophat
                assert np.abs(xe - oi["xhatend"]).max() < tz * (1 + np.abs(oi["xhatend"]).max()), (k, i)
                assert np.abs(gi["Yhat0"][i] + m.Yop - oi["Yhat"]).max() < tz * (1 + np.abs(oi["Yhat"]).max()), (k, i)
    assert n_active > 10
    print("all features worst", worst, "active", n_active, b.launch_info())


def _gpu_model2_with_disturbance(N):
    """The plant of test/3_test_predictive_control.jl:466-467 (see tests/test_oracle_linmpc.py) as a batch of N copies."""
    import mpc_b200
    from oracle.linmpc import zoh_first_order
    A1, B1, C1 = zoh_first_order(2, 10, 3.0)
    A2, B2, C2 = zoh_first_order(0.1, 7, 3.0)
    A = np.diag([A1[0, 0], A2[0, 0]])
    Bu, Bd = np.array([[B1[0, 0]], [0.0]]), np.array([[0.0], [B2[0, 0]]])
    C = np.array([[C1[0, 0], C2[0, 0]]])
    return mpc_b200.LinModel(A, Bu, C, Bd=Bd, Dd=np.zeros((1, 1)), Ts=3.0, N=N, uop=[25], dop=[30], yop=[50]), (A, Bu, C, Bd)


@pytest.mark.parametrize("kw,wmin,wmax,steps", [
    (dict(Wy=[[1]]), 36, 75, [(0, "Yhat", 36), (100, "Yhat", 75)]),
    (dict(Wu=[[1]]), 4, 20, [(0, "U", 4), (100, "U", 20)]),
    (dict(Wd=[[1]], Wy=[[1]]), 56, 95, [(0, "Yhat", 56 - 30), (100, "Yhat", 95 - 30)]),
    (dict(Wr=[[1]], Wy=[[1]]), 52, 175, [(21, "Yhat", 52 - 21), (100, "Yhat", 175 - 100)]),
])
def test_custom_linear_constraints_known_answers_gpu(kw, wmin, wmax, steps):
    """The eight known answers of test/3_test_predictive_control.jl:468-496 through the CUDA path (hard custom rows,
    Cwt = Inf, Hp = Hc = 50: n = 50, the CTA-team kernel), and the same steps against the oracle at the parity tolerance."""
    import mpc_b200
    from oracle.linmpc import LinModel as OLinModel, LinMPC as OLinMPC
    gm, (A, Bu, C, Bd) = _gpu_model2_with_disturbance(3)
    g = mpc_b200.LinMPC(gm, Nwt=[0], Cwt=np.inf, Hp=50, Hc=50, **kw)
    g.setconstraint(wmin=[wmin], wmax=[wmax])
    o = OLinMPC(OLinModel(A, Bu, C, Bd=Bd, Dd=np.zeros((1, 1)), Ts=3.0, uop=[25], dop=[30], yop=[50]), Nwt=[0],
                Cwt=np.inf, Hp=50, Hc=50, **kw)
    o.setconstraint(wmin=[wmin], wmax=[wmax])
    g.preparestate([[50]] * 3, [[30]] * 3)
    o.preparestate([50], [30])
    for ry, key, expect in steps:
        ug = g.moveinput([[ry]] * 3, [[30]] * 3)
        uo = o.moveinput([ry], [30])
        assert (g.batch.status == 0).all()
        info = g.getinfo()
        assert np.allclose(info[key], expect, atol=1e-1), (kw, ry, info[key][0][:5])
        assert np.abs(g.Ztilde[0] - o.Ztilde).max() < 5e-6 * (1 + np.abs(o.Ztilde).max())
        assert np.abs(ug[0] - uo).max() < 5e-6 * (1 + np.abs(uo).max())
        assert np.array_equal(g.Ztilde[0], g.Ztilde[2])


def test_custom_linear_constraints_warp_kernel_soft_rows_vs_oracle():
    """Soft custom rows mixing outputs, inputs and setpoints on the C1 shape (n = 11: warp kernel), closed loop vs the oracle."""
    import mpc_b200
    from mpc_b200 import workloads
    from oracle.linmpc import LinModel as OLinModel, LinMPC as OLinMPC
    N, steps = 6, 12
    model, rng = workloads.random_plants(N, 4, 2, 2, seed=17)
    Wy, Wu, Wr = [[1.0, -0.5]], [[0.3, 0.2]], [[-0.2, 0.0]]
    kw = dict(Hp=20, Hc=5, Cwt=1e4, Wy=Wy, Wu=Wu, Wr=Wr)
    g = mpc_b200.LinMPC(model, **kw).setconstraint(umin=[-1, -1], umax=[1, 1], wmin=[-0.4], wmax=[0.5], c_wmax=[0.7])
    os_, plants = [], []
    for i in range(N):
        o = OLinMPC(OLinModel(model.A[i], model.Bu[i], model.C[i]), **kw)
        o.setconstraint(umin=[-1, -1], umax=[1, 1], wmin=[-0.4], wmax=[0.5], c_wmax=[0.7])
        os_.append(o)
        plants.append(OLinModel(model.A[i], model.Bu[i], model.C[i]))
    ry = workloads.setpoints(rng, N, 2, steps, period=6)
    worst, nact = 0.0, 0
    for k in range(steps):
        y = np.stack([p.evaloutput() for p in plants])
        g.preparestate(y)
        ug = g.moveinput(ry[k])
        assert (g.batch.status == 0).all()
        nact += int((g.batch.iters > 0).sum())
        for i, o in enumerate(os_):
            o.preparestate(y[i])
            uo = o.moveinput(ry[k, i])
            e = np.abs(g.Ztilde[i] - o.Ztilde).max() / (1 + np.abs(o.Ztilde).max())
            assert e < 5e-6, (k, i, e)
            worst = max(worst, e)
            o.updatestate(uo, y[i])
            plants[i].updatestate(uo)
        g.updatestate(ug, y)
    assert nact > 10 and g.batch.launch_info()["team"] == 32
    print("custom rows, warp kernel: worst", worst, "active solves", nact)


def test_multiple_shooting_vs_oracle_ms():
    """LinMPC(transcription = MultipleShooting()) through the CUDA path: the condensed solve + bmpc_get_states against the
    oracle's ACTUAL MultipleShooting QP (equality constrained, oracle/linmpc_ms.py): Z̃ = [ΔU; X̂0; ε], u, J over a closed loop
    with hard input / increment bounds and soft output bounds; then the reference's known answers (test/3:120-127)."""
    import mpc_b200
    from mpc_b200 import workloads
    from oracle.linmpc import LinModel as OLinModel, zoh_first_order
    from oracle.linmpc_ms import LinMPCMultipleShooting
    N, steps = 5, 10
    model, rng = workloads.random_plants(N, 3, 2, 2, seed=29)
    kw = dict(Hp=9, Hc=3, Cwt=1e4, Lwt=[0.05, 0.0])
    cons = dict(umin=[-1, -1], umax=[1, 1], dumin=[-0.6, -0.6], dumax=[0.6, 0.6], ymax=[0.7, 0.7])
    g = mpc_b200.LinMPC(model, transcription="MultipleShooting", **kw).setconstraint(**cons)
    os_ = [LinMPCMultipleShooting(OLinModel(model.A[i], model.Bu[i], model.C[i]), **kw).setconstraint(**cons) for i in range(N)]
    plants = [OLinModel(model.A[i], model.Bu[i], model.C[i]) for i in range(N)]
    ry = workloads.setpoints(rng, N, 2, steps, period=5)
    worst, nact = 0.0, 0
    for k in range(steps):
        y = np.stack([p.evaloutput() for p in plants])
        g.preparestate(y)
        ug = g.moveinput(ry[k])
        Zg = g.Ztilde
        info = g.getinfo()
        assert Zg.shape == (N, os_[0].n) and (g.batch.status == 0).all()
        nact += int((g.batch.iters > 0).sum())
        for i, o in enumerate(os_):
            o.preparestate(y[i])
            uo = o.moveinput(ry[k, i])
            e = np.abs(Zg[i] - o.Ztilde).max() / (1 + np.abs(o.Ztilde).max())
            assert e < 5e-6 and np.abs(ug[i] - uo).max() < 5e-6, (k, i, e)
            io = o.getinfo()
            assert abs(info["J"][i] - io["J"]) < 1e-7 * (1 + abs(io["J"]))
            assert np.abs(info["X0"][i] - io["X0"]).max() < 5e-6 * (1 + np.abs(io["X0"]).max())
            worst = max(worst, e)
            o.updatestate(uo, y[i])
            plants[i].updatestate(uo)
        g.updatestate(ug, y)
    assert nact > 5
    print("MultipleShooting vs oracle MS: worst", worst, "active solves", nact)
    # test/3_test_predictive_control.jl:120-127
    A, B, C = zoh_first_order(5, 2, 3.0)
    g5 = mpc_b200.LinMPC(mpc_b200.LinModel(A, B, C, Ts=3.0, yop=[10]), Nwt=[0], Hp=1000, Hc=1, transcription="multipleshooting")
    g5.preparestate([[10]])
    u = g5.moveinput([[15]])
    assert abs(u[0, 0] - 1) < 1e-2 and abs(g5.getinfo()["Yhat"][0, -1] - 15) < 1e-2
    assert g5.Ztilde.shape == (1, 1 + 2 * 1000 + 1)


def test_dense_L_Hp_and_N_Hc_weights_vs_oracle():
    """Dense (non-diagonal, symmetric positive semidefinite) L_Hp, Ñ_Hc and M_Hp weight matrices (ControllerWeights,
    src/controller/construct.jl:45-93) with an input-setpoint trajectory R̂u, operating points and active constraints:
    route B carries Ñ_Hc inside H̃, the kernel forms L_Hp Cu and M_Hp Cy itself.  GPU vs oracle, closed loop."""
    from helpers import random_plant
    rng = np.random.default_rng(77)
    N, Hp, Hc, nu, ny = 4, 6, 3, 2, 2
    mpcs, plants = [], []
    psd = lambda n, s: (lambda Q: s * (Q @ Q.T) / n + 0.05 * np.eye(n))(rng.standard_normal((n, n)))
    for i in range(N):
        p = random_plant(rng, nx=3, nu=nu, ny=ny)
        m = LinModel(p.A, p.Bu, p.C, uop=[0.2, -0.1], yop=[1.0, 0.5])
        mpc = LinMPC(m, Hp=Hp, Hc=Hc, M_Hp=psd(ny * Hp, 1.0), N_Hc=psd(nu * Hc, 0.1), L_Hp=psd(nu * Hp, 0.3), Cwt=1e4)
        mpc.setconstraint(umin=[-0.8, -0.8], umax=[0.8, 0.8], ymax=[1.6, 1.2])
        mpcs.append(mpc)
        plants.append(LinModel(p.A, p.Bu, p.C, uop=[0.2, -0.1], yop=[1.0, 0.5]))
    b = batch_from_oracle(mpcs)
    worst, nact = 0.0, 0
    for k in range(10):
        ry = np.array([1.0, 0.5]) + rng.choice([-1.0, 1.0], (N, ny))
        Ru = np.tile([0.2, -0.1], Hp) + 0.3 * rng.standard_normal((N, nu * Hp))
        ys = [p.evaloutput() for p in plants]
        for m, y in zip(mpcs, ys):
            m.preparestate(y)
        b.lastu0[:] = np.stack([m.lastu0 for m in mpcs])
        ug = b.step(np.stack([m.estim.xhat0 for m in mpcs]), ry=ry, Rhat_u=Ru).copy()
        assert (b.status == 0).all()
        nact += int((b.iters > 0).sum())
        for i, m in enumerate(mpcs):
            u = m.moveinput(ry[i], Rhat_u=Ru[i])
            e = np.abs(b.Ztilde[i] - m.Ztilde).max() / (1 + np.abs(m.Ztilde).max())
            assert e < 5e-6 and abs(b.J[i] - m.getinfo()["J"]) < 1e-8 * (1 + abs(m.getinfo()["J"])), (k, i, e)
            worst = max(worst, e)
            m.updatestate(u, ys[i])
            plants[i].updatestate(u)
    assert nact > 5
    print("dense L_Hp / N_Hc / M_Hp: worst", worst, "active solves", nact)
