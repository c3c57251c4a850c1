"""Oracle MultipleShooting MHE (oracle/mhe_ms.py) pinned to the reference's inline known answers
(test/2_test_state_estim.jl:1126-1139 and :1718-1726) and cross-checked against the SingleShooting oracle: the two
transcriptions of the same estimation problem must give the same arrival state, process noises, state trajectory,
estimate and objective.  The MS problem is solved as an equality-constrained QP by a generic null-space method, so the
agreement is an independent check of the condensation the CUDA path uses."""
import numpy as np
import pytest

from oracle import qp
from oracle.linmpc import LinModel
from oracle.mhe import MovingHorizonEstimator
from oracle.mhe_ms import MovingHorizonEstimatorMS


def setup_linmodel():
    """test/0_test_module.jl:3-8 + test/2:1036-1037: sys = [1.9/(1800s+1) x3; -0.74, 0.74, -0.74 /(800s+1)], Ts = 400,
    i_u = [1, 2], i_d = [3], zero-order hold (a minimal realisation: one state per output)."""
    Ts, a1, a2 = 400.0, np.exp(-400.0 / 1800.0), np.exp(-400.0 / 800.0)
    A = np.diag([a1, a2])
    Bu = np.array([[1.9 * (1 - a1), 1.9 * (1 - a1)], [-0.74 * (1 - a2), 0.74 * (1 - a2)]])
    Bd = np.array([[1.9 * (1 - a1)], [-0.74 * (1 - a2)]])
    return LinModel(A, Bu, np.eye(2), Bd=Bd, Ts=Ts, uop=[10, 50], yop=[50, 30], dop=[5])


@pytest.mark.parametrize("direct", [True, False])
def test_mhe_ms_known_answer_outputs(direct):
    """test/2:1126-1139: mhe5 / mhe6 = MovingHorizonEstimator(linmodel, He=2, direct, transcription=MultipleShooting()):
    after 40 periods at ym = [51, 32] the estimated output is [51, 32] (atol 1e-3)."""
    mhe = MovingHorizonEstimatorMS(setup_linmodel(), He=2, direct=direct)
    for _ in range(40):
        mhe.preparestate([51, 32], [5])
        mhe.updatestate([10, 50], [51, 32], [5])
        assert mhe.last_qp["status"] == qp.OPTIMAL
    mhe.preparestate([51, 32], [5])
    assert mhe.evaloutput([5]) == pytest.approx([51, 32], abs=1e-3)


def test_mhe_ms_known_answer_operating_points():
    """test/2:1718-1726: A = 0.5, Bu = 0.3, C = 1, xop = fop = 3, uop = 2, yop = 50, He = 5, nint_ym = 0, direct = false,
    MultipleShooting: updatestate!(mhe3, [2 + 1], [50]) ~ [3 + 0.3]."""
    m = LinModel([[0.5]], [[0.3]], [[1.0]], Ts=10.0, uop=[2.0], yop=[50.0], xop=[3.0], fop=[3.0])
    mhe = MovingHorizonEstimatorMS(m, He=5, nint_ym=0, direct=False)
    x = mhe.updatestate([3.0], [50.0])
    assert x == pytest.approx([3.3], abs=1e-3)


def _plant(rng, nx=3, nu=2, ny=2, nd=1, same_fop=False):
    A = rng.standard_normal((nx, nx))
    A *= 0.85 / np.abs(np.linalg.eigvals(A)).max()
    xop = rng.standard_normal(nx) * 0.1
    return LinModel(A, rng.standard_normal((nx, nu)), rng.standard_normal((ny, nx)), Bd=rng.standard_normal((nx, nd)),
                    Dd=0.1 * rng.standard_normal((ny, nd)), uop=[1, 2], yop=[5, 3], dop=[2], xop=xop,
                    fop=xop.copy() if same_fop else rng.standard_normal(nx) * 0.1)


@pytest.mark.parametrize("direct", [True, False])
@pytest.mark.parametrize("case", ["free", "hard", "soft"])
def test_mhe_ms_equals_single_shooting(direct, case):
    rng = np.random.default_rng(11 + (1 if direct else 0))
    for trial in range(2):
        # direct = false with f̂op != x̂op: the reference's SingleShooting `B` leaves the LAST measurement block without
        # its (f̂op - x̂op) term once the window is full (`row_end = He-2`, mhe/transcription.jl:243-250; restated as is in
        # oracle/mhe.py), so the two transcriptions only coincide for f̂op = x̂op there -- which is also all the reference's
        # own tests use in that mode (test/2:1718-1719)
        model = _plant(rng, same_fop=not direct)
        kw = dict(He=4, direct=direct, Cwt=(1e4 if case == "soft" else np.inf))
        ss, ms = MovingHorizonEstimator(model, **kw), MovingHorizonEstimatorMS(model, **kw)
        nxh = ss.nxhat
        if case != "free":
            cons = dict(xhatmin=[-0.8] * nxh, xhatmax=[0.9] * nxh, whatmin=[-0.15] * nxh, whatmax=[0.15] * nxh,
                        vhatmin=[-0.4, -0.4], vhatmax=[0.4, 0.4])
            if case == "hard":  # wide enough to stay feasible under the 0.5-sigma measurement noise, tight enough to bind
                cons.update(whatmin=[-0.3] * nxh, whatmax=[0.3] * nxh, vhatmin=[-1.5, -1.5], vhatmax=[1.5, 1.5],
                            xhatmin=[-1.5] * nxh, xhatmax=[1.5] * nxh)
            if case == "soft":
                cons.update(c_xhatmin=[0.5] * nxh, c_xhatmax=[0.5] * nxh, c_vhatmin=[1, 1], c_vhatmax=[1, 1])
            ss.setconstraint(**cons), ms.setconstraint(**cons)
        nact, feasible = 0, True
        for k in range(10):                        # growing window (k < He), then moving
            y = np.array([5.0, 3.0]) + 0.5 * rng.standard_normal(2)
            d, u = [2.0 + 0.1 * rng.standard_normal()], np.array([1.0, 2.0]) + 0.3 * rng.standard_normal(2)
            xs, xm = ss.preparestate(y, d).copy(), ms.preparestate(y, d).copy()
            if direct:
                feasible = feasible and _check(ss, ms, xs, xm, (trial, k))
                nact += int((ss.last_qp["lam"] is not None) and np.any(np.asarray(ss.last_qp["lam"]) > 1e-8))
            if not feasible:
                break  # after an infeasible window the fallbacks (shifted warm starts) differ by transcription
            xs, xm = ss.updatestate(u, y, d).copy(), ms.updatestate(u, y, d).copy()
            if not direct:
                feasible = feasible and _check(ss, ms, xs, xm, (trial, k))
                nact += int((ss.last_qp["lam"] is not None) and np.any(np.asarray(ss.last_qp["lam"]) > 1e-8))
        if case != "free":
            assert feasible and nact > 0, "the bounds never became active: the case does not exercise the constraints"


def _check(ss, ms, xs, xm, where):
    if ss.last_qp["status"] != qp.OPTIMAL:       # (infeasible windows of the hard case: both must say so)
        assert ms.last_qp["status"] == ss.last_qp["status"], where
        return False
    assert ms.last_qp["status"] == qp.OPTIMAL, where
    neps, nxh, He, Nk = ss.neps, ss.nxhat, ss.He, ss.Nk
    tol = 2e-7
    assert np.abs(xs - xm).max() < tol, where
    nxt = neps + nxh
    assert np.abs(ms.Ztilde[:nxt] - ss.Ztilde[:nxt]).max() < tol, where                       # ε, arrival state
    assert np.abs(ms.Ztilde[nxt + nxh * He:] - ss.Ztilde[nxt:]).max() < tol, where             # Ŵ
    assert np.abs(ms.Ztilde[nxt:nxt + nxh * Nk] - ss.X0[:nxh * Nk]).max() < tol, where         # X̂0 block = SS trajectory
    assert np.all(ms.Ztilde[nxt + nxh * Nk:nxt + nxh * He] == 0.0), where                      # fill0unused!
    assert abs(ms.Jval - ss.Jval) < 1e-7 * (1 + abs(ss.Jval)), where
    assert np.abs(ms.Vhat - ss.Vhat).max() < tol, where
    return True
