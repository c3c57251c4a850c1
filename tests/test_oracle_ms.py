"""Oracle MultipleShooting LinMPC (oracle/linmpc_ms.py) pinned to the reference's inline known answers
(test/3_test_predictive_control.jl:120-127 and :570-579) and cross-checked against the SingleShooting oracle:
the two transcriptions of the same optimal-control problem must give the same ΔU, ε and objective, and the X̂0 block
of the MS decision vector must be the state trajectory the SS prediction implies."""
import numpy as np
import pytest

from oracle import qp
from oracle.linmpc import LinModel, LinMPC, zoh_first_order
from oracle.linmpc_ms import LinMPCMultipleShooting
from oracle.mhe import KalmanFilter
from helpers import random_plant


def test_ms_known_answer_moveinput():
    """test/3:120-127: mpc5 = LinMPC(linmodel, Hp=1000, Hc=1, transcription=MultipleShooting()): u ~ 1, Ŷ[end] ~ 15."""
    m = LinModel(*zoh_first_order(5, 2, 3.0), Ts=3.0, yop=[10])
    mpc = LinMPCMultipleShooting(m, Nwt=[0], Hp=1000, Hc=1)   # (linmodel of the test file: tf(5,[2,1]), Ts=3, yop=10)
    mpc.preparestate([10])
    u = mpc.moveinput([15])
    assert mpc.last_status == qp.OPTIMAL
    info = mpc.getinfo()
    assert u == pytest.approx([1], abs=1e-2) and info["u"] == pytest.approx([1], abs=1e-2)
    assert info["Yhat"][-1] == pytest.approx(15, abs=1e-2)


def test_ms_known_answer_setmodel():
    """test/3:570-579: KalmanFilter + MultipleShooting, u ~ 3 then, after setmodel! to twice the gain, u ~ 4."""
    mpc = LinMPCMultipleShooting(KalmanFilter(LinModel(*zoh_first_order(5, 2, 3.0), Ts=3.0)), Nwt=[0], Hp=1000, Hc=1)
    mpc.preparestate([0])
    u = mpc.moveinput([15])
    assert u == pytest.approx([3], abs=1e-2)
    mpc.setmodel(LinModel(*zoh_first_order(10, 2, 3.0), Ts=3.0))
    u = mpc.moveinput([40])
    assert u == pytest.approx([4], abs=1e-2)


@pytest.mark.parametrize("case", ["soft", "hard_terminal"])
def test_ms_equals_single_shooting(case):
    rng = np.random.default_rng(8)
    for trial in range(3):
        p = random_plant(rng, nx=3, nu=2, ny=2)
        if case == "soft":
            kw = dict(Hp=8, Hc=3, Cwt=1e4, Lwt=[0.1, 0.05])
            cons = dict(umin=[-1, -1], umax=[1, 1], ymax=[0.6, 0.7], dumin=[-0.5, -0.5], dumax=[0.5, 0.5], c_dumax=[0.2, 0.2])
        else:
            kw = dict(Hp=7, Hc=[1, 2, 2], Cwt=np.inf)
            cons = dict(umin=[-0.7, -0.7], umax=[0.7, 0.7], xhatmin=[-2.5] * 5, xhatmax=[2.5] * 5, dumin=[-0.4, -0.4], dumax=[0.4, 0.4])
        ss = LinMPC(LinModel(p.A, p.Bu, p.C), **kw).setconstraint(**cons)
        ms = LinMPCMultipleShooting(LinModel(p.A, p.Bu, p.C), **kw).setconstraint(**cons)
        plant = LinModel(p.A, p.Bu, p.C)
        nact = 0
        for k in range(8):
            ry = rng.choice([-1.0, 1.0], 2)
            y = plant.evaloutput()
            ss.preparestate(y), ms.preparestate(y)
            us, um = ss.moveinput(ry), ms.moveinput(ry)
            assert ss.last_status == qp.OPTIMAL and ms.last_status == qp.OPTIMAL
            nDU = ms.nDU
            assert np.abs(ms.Ztilde[:nDU] - ss.Ztilde[:nDU]).max() < 1e-7, (trial, k)
            assert np.abs(um - us).max() < 1e-7
            i_s, i_m = ss.getinfo(), ms.getinfo()
            assert abs(i_s["J"] - i_m["J"]) < 1e-8 * (1 + abs(i_s["J"]))
            assert np.abs(i_s["Yhat"] - i_m["Yhat"]).max() < 1e-7 and np.abs(i_s["xhatend"] - i_m["xhatend"]).max() < 1e-7
            # X̂0 block = the state recursion driven by the optimal inputs
            x = ms.estim.xhat0.copy()
            U0 = i_m["U"] - ms.Uop
            for j in range(ms.Hp):
                x = ms.estim.Ahat @ x + ms.estim.Buhat @ U0[2 * j:2 * j + 2] + ms.estim.fophat - ms.estim.xophat
                assert np.abs(i_m["X0"][x.size * j:x.size * (j + 1)] - x).max() < 1e-7
            nact += int(np.abs(ss.last_qp["lam"]).max() > 1e-9) if ss.last_qp["lam"] is not None and len(ss.last_qp["lam"]) else 0
            ss.updatestate(us, y), ms.updatestate(um, y)
            plant.updatestate(us)
        assert nact > 0
