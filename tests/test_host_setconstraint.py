"""Host-side logic of the mirror's ``setconstraint`` (modelpredictivecontrol.jl_b200/linmpc.py) on the CPU: the method is run
against a stand-in for the controller object (no device, no library call -- `_push`, which hands the compiled bounds to
the C ABI, is replaced by a recorder), and must store exactly what the oracle's restatement of ``setconstraint!``
(src/controller/construct.jl:324-559) stores for the same keywords, with the same error behaviour
(test/3_test_predictive_control.jl:259-389)."""
import types

import numpy as np
import pytest

import mpc_b200
from mpc_b200.host import expand_softness
from oracle.linmpc import LinModel as OLinModel, LinMPC as OLinMPC, zoh_first_order


def _stand_in(N, nu, ny, nx, Hp, Hc, nw, neps=1):
    inf = np.inf
    con = dict(U0min=np.full((N, nu * Hp), -inf), U0max=np.full((N, nu * Hp), inf), DUmin=np.full((N, nu * Hc), -inf),
               DUmax=np.full((N, nu * Hc), inf), Y0min=np.full((N, ny * Hp), -inf), Y0max=np.full((N, ny * Hp), inf),
               Wmin=np.full((N, nw * (Hp + 1)), -inf), Wmax=np.full((N, nw * (Hp + 1)), inf),
               xhat0min=np.full((N, nx), -inf), xhat0max=np.full((N, nx), inf))
    o = types.SimpleNamespace(
        model=types.SimpleNamespace(N=N, nu=nu, ny=ny), Hp=Hp, Hc=Hc, nw=nw, con=con,
        estim=types.SimpleNamespace(nxhat=nx, xophat=np.zeros((N, nx))), Uop=np.zeros((N, nu * Hp)), Yop=np.zeros((N, ny * Hp)),
        batch=types.SimpleNamespace(neps=neps), _solved=False, pushes=0,
        soft=dict(C_umin=np.zeros(nu * Hp), C_umax=np.zeros(nu * Hp), C_dumin=np.zeros(nu * Hc), C_dumax=np.zeros(nu * Hc),
                  C_ymin=np.ones(ny * Hp), C_ymax=np.ones(ny * Hp), c_xmin=np.ones(nx), c_xmax=np.ones(nx)),
        soft_w=dict(C_wmin=np.ones(nw * (Hp + 1)), C_wmax=np.ones(nw * (Hp + 1))))
    o._push = lambda: setattr(o, "pushes", o.pushes + 1)
    o.setconstraint = lambda **kw: mpc_b200.LinMPC.setconstraint(o, **kw)
    return o


def test_expand_softness():
    assert expand_softness(None, None, 2, 3, "umin") is None
    assert np.array_equal(expand_softness([1, 2], None, 2, 3, "umin"), [1, 2, 1, 2, 1, 2])
    assert np.array_equal(expand_softness([9, 9], np.arange(6.0), 2, 3, "umin"), np.arange(6.0))  # the whole-horizon form wins
    for bad in (dict(small=[1, 2, 3], big=None), dict(small=None, big=[1, 2, 3]), dict(small=[-1, 0], big=None)):
        with pytest.raises(ValueError):
            expand_softness(bad["small"], bad["big"], 2, 3, "umin")


def test_mirror_setconstraint_matches_oracle_and_reference_errors():
    N, nu, ny, nx, Hp, Hc, nw = 3, 1, 1, 2, 50, 5, 1
    g = _stand_in(N, nu, ny, nx, Hp, Hc, nw)
    o = OLinMPC(OLinModel(*zoh_first_order(2, 10, 3.0), Ts=3.0), Hp=Hp, Hc=Hc, Wr=[[1]])
    r50, r5, r51 = np.arange(1, 51.0), np.arange(1, 6.0), np.arange(1, 52.0)
    calls = [dict(umin=[-3], umax=[4]), dict(dumin=[-1], dumax=[2]), dict(ymin=[-6], ymax=[55]), dict(wmin=[-7], wmax=[75]),
             dict(xhatmin=[-21, -22], xhatmax=[21, 22]),
             dict(Umin=-r50 - 1, Umax=r50 + 1), dict(DUmin=-r5 - 2, DUmax=r5 + 2), dict(Ymin=-r50 - 3, Ymax=r50 + 3),
             dict(Wmin=-r51 - 4, Wmax=r51 + 4),
             dict(c_umin=[0.01], c_umax=[0.03]), dict(c_dumin=[0.05], c_dumax=[0.07]), dict(c_ymin=[1.0], c_ymax=[1.02]),
             dict(c_wmin=[2.0], c_wmax=[2.02]), dict(c_xhatmin=[0.21, 0.22], c_xhatmax=[0.31, 0.32]),
             dict(C_umin=r50 + 5, C_umax=r50 + 5), dict(C_dumin=r5 + 6, C_dumax=r5 + 6), dict(C_ymin=r50 + 7, C_ymax=r50 + 7),
             dict(C_wmin=r51 + 8, C_wmax=r51 + 8)]
    pairs = [("U0min", "U0min"), ("U0max", "U0max"), ("DUmin", "DUmin"), ("DUmax", "DUmax"), ("Y0min", "Y0min"), ("Y0max", "Y0max"),
             ("Wmin", "Wmin"), ("Wmax", "Wmax"), ("xhat0min", "xhat0min"), ("xhat0max", "xhat0max")]
    for n, kw in enumerate(calls):
        g.setconstraint(**kw)
        o.setconstraint(**kw)
        assert g.pushes == n + 1
        for kg, ko in pairs:
            for i in range(N):  # every instance carries the broadcast bound
                assert np.array_equal(g.con[kg][i], getattr(o.con, ko)), (kw, kg)
        for k in ("C_umin", "C_umax", "C_dumin", "C_dumax", "C_ymin", "C_ymax", "c_xmin", "c_xmax"):
            assert np.array_equal(g.soft[k], getattr(o.con, k)), (kw, k)
        for k in ("C_wmin", "C_wmax"):
            assert np.array_equal(g.soft_w[k], getattr(o.con, k)), (kw, k)
    # per-instance bounds: (N, len) arrays are taken row by row
    g.setconstraint(umin=np.array([[-1.0], [-2.0], [-3.0]]))
    assert np.array_equal(g.con["U0min"][:, 0], [-1, -2, -3]) and np.array_equal(g.con["U0min"][:, -1], [-1, -2, -3])
    # errors (DimensionMismatch / negative softness -> ValueError, nothing stored)
    before = {k: v.copy() for k, v in g.soft.items()}
    for kw in ("umin", "umax", "dumin", "dumax", "ymin", "ymax", "wmin", "wmax",
               "c_umin", "c_umax", "c_dumin", "c_dumax", "c_ymin", "c_ymax", "c_wmin", "c_wmax"):
        with pytest.raises(ValueError):
            g.setconstraint(**{kw: [0, 0, 0]})
    for kw in ("c_umin", "c_umax", "c_dumin", "c_dumax", "c_ymin", "c_ymax", "c_wmin", "c_wmax", "C_umin"):
        with pytest.raises(ValueError):
            g.setconstraint(**{kw: -np.ones(50 if kw == "C_umin" else 1)})
    assert all(np.array_equal(before[k], g.soft[k]) for k in before)
    g._solved = True
    with pytest.raises(RuntimeError):  # softness is frozen after the first moveinput!
        g.setconstraint(c_umin=[1], c_umax=[1])
    g.setconstraint(umin=[-9])          # values may still change
    hard = _stand_in(N, nu, ny, nx, Hp, Hc, nw, neps=0)
    for kw in ("c_umin", "c_umax", "c_dumin", "c_dumax", "c_ymin", "c_ymax", "C_ymax"):
        with pytest.raises(ValueError):  # ArgumentError: no slack variable (Cwt = Inf)
            hard.setconstraint(**{kw: np.ones(50 if kw == "C_ymax" else 1)})
    nowt = _stand_in(N, nu, ny, nx, Hp, Hc, 0)
    with pytest.raises(ValueError):
        nowt.setconstraint(wmin=[0])


def test_mhe_mirror_setconstraint_on_stand_in(monkeypatch):
    """Same for the estimator mirror (modelpredictivecontrol.jl_b200/mhe.py::setconstraint): the C-ABI call is replaced by
    a recorder; bounds go out in deviation form per instance, softness as [min; max] pairs; sizes, negative weights, the
    post-solve freeze and Cwt = Inf raise as in the reference (test/2_test_state_estim.jl:1452-1489)."""
    from mpc_b200 import _lib, mhe as gmhe
    calls = []

    class FakeLib:
        def bmhe_set_constraints(self, h, *ptrs):
            calls.append(ptrs)
            return 0
    monkeypatch.setattr(_lib, "lib", lambda: FakeLib())
    N, nxh, nym = 3, 2, 2
    inf = np.inf
    o = types.SimpleNamespace(
        model=types.SimpleNamespace(N=N), nxhat=nxh, nym=nym, xophat=np.tile([1.0, 2.0], (N, 1)), neps=1, _solved=False, _h=None,
        con=dict(xmin=np.full((N, nxh), -inf), xmax=np.full((N, nxh), inf), wmin=np.full((N, nxh), -inf),
                 wmax=np.full((N, nxh), inf), vmin=np.full((N, nym), -inf), vmax=np.full((N, nym), inf)),
        soft=dict(c_x=np.zeros(2 * nxh), c_w=np.zeros(2 * nxh), c_v=np.zeros(2 * nym)))
    sc = lambda **kw: gmhe.MovingHorizonEstimator.setconstraint(o, **kw)
    sc(xhatmin=[-51, -52], xhatmax=[53, 54])
    assert np.array_equal(o.con["xmin"], np.tile([-52.0, -54.0], (N, 1))) and np.array_equal(o.con["xmax"], np.tile([52.0, 52.0], (N, 1)))
    sc(whatmin=[-55, -56], whatmax=[57, 58], vhatmin=[-59, -60], vhatmax=[61, 62])
    assert np.array_equal(o.con["wmin"][1], [-55, -56]) and np.array_equal(o.con["vmax"][2], [61, 62])
    sc(c_xhatmin=[0.01, 0.02], c_xhatmax=[0.03, 0.04], c_whatmin=[0.05, 0.06], c_whatmax=[0.07, 0.08],
       c_vhatmin=[0.09, 0.10], c_vhatmax=[0.11, 0.12])
    assert np.allclose(o.soft["c_x"], [0.01, 0.02, 0.03, 0.04]) and np.allclose(o.soft["c_w"], [0.05, 0.06, 0.07, 0.08])
    assert np.allclose(o.soft["c_v"], [0.09, 0.10, 0.11, 0.12]) and len(calls) == 3
    before = {k: v.copy() for k, v in o.soft.items()}
    for kw in ("xhatmin", "xhatmax", "whatmin", "whatmax", "vhatmin", "vhatmax",
               "c_xhatmin", "c_xhatmax", "c_whatmin", "c_whatmax", "c_vhatmin", "c_vhatmax"):
        with pytest.raises(ValueError):
            sc(**{kw: [1.0]})
    with pytest.raises(ValueError):
        sc(c_xhatmin=[-1, 0], c_vhatmax=[5, 5])
    assert all(np.array_equal(before[k], o.soft[k]) for k in before) and len(calls) == 3
    o._solved = True
    with pytest.raises(RuntimeError):
        sc(c_xhatmin=[100, 100])
    o._solved, o.neps = False, 0
    with pytest.raises(ValueError):
        sc(c_whatmax=[1, 1])
