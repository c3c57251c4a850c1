"""BASELINE.json configs[2], [3], [4] at their full (per-GPU) sizes through the C ABI: size-independent properties
that need no oracle -- every instance solves, bounds are honoured, results do not depend on the instance order
(bitwise), duplicated instances agree bitwise -- plus an oracle spot check on a few sampled instances
(tolerance 5e-6 relative on Z̃ when the interior-point solver ran, 1e-9 otherwise; J 1e-7 relative: at a
constrained optimum J is first-order sensitive to Z̃ -- dJ/dε = 2 Cwt ε ≈ 64 on these problems -- so a Z̃ agreement
of 1e-8 gives 4e-7 absolute on J ≈ 20, measured with tools/studies/c4_jcheck.py; both solutions satisfy every
constraint to 1e-14)."""
import numpy as np
import pytest

from oracle import qp
from oracle.linmpc import LinModel as OLinModel, LinMPC as OLinMPC, ManualEstimator as OManual
from oracle.mhe import MovingHorizonEstimator as OMHE

pytestmark = pytest.mark.gpu


def _linmpc_case(name, N, con, sample, periods=2):
    import mpc_b200
    from mpc_b200 import workloads
    _, nx, nu, ny, Hp, Hc, seed = workloads.CONFIGS[name]
    model, rng = workloads.random_plants(N, nx, nu, ny, seed)
    model.A[1], model.Bu[1], model.C[1] = model.A[0], model.Bu[0], model.C[0]  # duplicate instance
    xh = rng.standard_normal((periods, N, nx + ny)) * 0.3
    ry = rng.choice([-1.0, 1.0], (periods, N, ny))
    xh[:, 1], ry[:, 1] = xh[:, 0], ry[:, 0]

    def run(perm):
        sub = mpc_b200.LinModel(model.A[perm], model.Bu[perm], model.C[perm], N=N)
        mpc = mpc_b200.LinMPC(mpc_b200.ManualEstimator(sub), Hp=Hp, Hc=Hc, Cwt=1e5)
        mpc.setconstraint(**con(nu, ny))
        out = []
        for k in range(periods):  # the second period is warm-started from the first one's solution and multipliers
            mpc.estim.xhat0 = xh[k][perm].copy()
            mpc.moveinput(ry[k][perm])
            out.append((mpc.Ztilde.copy(), mpc.getinfo()))
        return mpc, out

    ident = np.arange(N)
    mpc, out = run(ident)
    c = con(nu, ny)
    for Z, info in out:
        assert (info["status"] == 0).all(), np.bincount(info["status"])
        assert 0 < info["iters"].max() < 50
        assert info["U"].max() <= c["umax"][0] + 1e-7 and info["U"].min() >= c["umin"][0] - 1e-7
        if "dumax" in c:
            assert np.abs(info["DU"]).max() <= c["dumax"][0] + 1e-7
        assert (info["Yhat"] - c["ymax"][0] - info["eps"][:, None]).max() <= 1e-6
        if "ymin" in c:
            assert (c["ymin"][0] - info["Yhat"] - info["eps"][:, None]).max() <= 1e-6
        assert (info["eps"] >= -1e-12).all()
        assert np.array_equal(Z[0], Z[1])
    perm = np.random.default_rng(0).permutation(N)
    _, out_p = run(perm)
    for (Z, _), (Zp, _) in zip(out, out_p):
        assert np.array_equal(Zp, Z[perm])
    # oracle spot check on sampled instances (both periods; the oracle solves each QP exactly)
    worst = 0.0
    for i in sample:
        om = OLinModel(model.A[i], model.Bu[i], model.C[i])
        o = OLinMPC(OManual(om), Hp=Hp, Hc=Hc, Cwt=1e5)
        o.setconstraint(**c)
        for k in range(periods):
            o.estim.xhat0 = xh[k, i].copy()
            o.moveinput(ry[k, i])
            assert o.last_status == qp.OPTIMAL
            Z, info = out[k]
            tol = 5e-6 if info["iters"][i] > 0 else 1e-9
            ez = np.abs(Z[i] - o.Ztilde).max() / (1 + np.abs(o.Ztilde).max())
            Jo = o.getinfo()["J"]
            assert ez < tol, (name, i, k, ez)
            assert abs(info["J"][i] - Jo) <= 1e-7 * (1 + abs(Jo)), (name, i, k)
            worst = max(worst, ez)
    print(name, "full size", N, "worst vs oracle", worst, mpc.batch.launch_info())


def test_full_size_c2_properties():
    """configs[2]: 65 536 controllers, 4x4 plants (nx = 8), Hp = 30, Hc = 10, hard u box + soft ymax."""
    _linmpc_case("C2", 65536, lambda nu, ny: dict(umin=[-1.0] * nu, umax=[1.0] * nu, ymax=[0.8] * ny),
                 sample=[0, 7, 4099, 65535])


def test_full_size_c4_shard_properties():
    """configs[4]: 8x8 plants (nx = 16), Hp = 50, Hc = 20, hard u and du boxes + soft ymin/ymax; one GPU's shard
    (16 384 / 8 = 2048 controllers)."""
    _linmpc_case("C4", 2048, lambda nu, ny: dict(umin=[-1.0] * nu, umax=[1.0] * nu, dumin=[-0.2] * nu, dumax=[0.2] * nu,
                                                 ymin=[-1.2] * ny, ymax=[0.8] * ny), sample=[0, 1023])


def test_full_size_c3_mhe_properties():
    """configs[3]: 8192 linear MHE, He = 15, 4x4 plants (nx = 8, nint_ym = 0), bounds on x̂, ŵ, v̂; growing window,
    first full window and two moving windows."""
    import mpc_b200
    from mpc_b200 import workloads
    N, He, nx, nu, ny = 8192, 15, 8, 4, 4
    model, rng = workloads.random_plants(N, nx, nu, ny, seed=3)
    model.A[1], model.Bu[1], model.C[1] = model.A[0], model.Bu[0], model.C[0]
    kw = dict(xhatmin=[-10] * nx, xhatmax=[10] * nx, whatmin=[-0.5] * nx, whatmax=[0.5] * nx, vhatmin=[-3] * ny,
              vhatmax=[3] * ny)
    mhe = mpc_b200.MovingHorizonEstimator(model, He=He, nint_ym=[0] * ny).setconstraint(**kw)
    plant = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
    steps = He + 2
    Us, Ys = [], []
    for k in range(steps):  # the data do not depend on the estimator: generate them once
        u = rng.choice([-1.0, 1.0], (N, nu))
        plant.x0 = plant.x0 + 0.5 * rng.standard_normal((N, nx)) / nx
        y = plant.evaloutput() + 0.7 * rng.standard_normal((N, ny))
        u[1], plant.x0[1], y[1] = u[0], plant.x0[0], y[0]
        Us.append(u); Ys.append(y)
        plant.updatestate(u)
    # pass 1 (no checks): find instances whose windows keep the interior-point solver busy and stay feasible
    busy, feas = np.zeros(N, int), np.ones(N, bool)
    for k in range(steps):
        mhe.preparestate(Ys[k])
        busy += mhe.iters > 0
        feas &= mhe.status == 0
        mhe.updatestate(Us[k])
    sample = [0] + [int(i) for i in np.argsort(-(busy * feas))[:2]]
    mhe.reset()
    os_ = {i: OMHE(OLinModel(model.A[i], model.Bu[i], model.C[i]), He=He, nint_ym=0).setconstraint(**kw) for i in sample}
    worst, nact, stat1, nact_sample = 0.0, 0, 0, 0
    for k in range(steps):
        u, y = Us[k], Ys[k]
        x = mhe.preparestate(y).copy()
        ok = mhe.status == 0
        # status 2: the window is infeasible by construction (sensor noise against |v̂| <= 3, process noise against
        # |ŵ| <= 0.5); status 1: a marginal window on which the IPM reaches its iteration cap with the iterate kept, the
        # reference's @warn branch (mhe/execute.jl:590-610) -- allowed for at most 1 % of the batch
        n1, n2 = int((mhe.status == 1).sum()), int((mhe.status == 2).sum())
        stat1 += n1
        assert np.isin(mhe.status, (0, 1, 2)).all() and n1 <= N // 100, (k, n1, n2)
        assert ok.mean() > 0.9, ok.mean()
        nz = nx * (1 + min(k + 1, He))
        W = mhe.Ztilde[ok][:, nx:nz]
        assert np.abs(W).max() <= 0.5 + 1e-7        # ŵ bounds
        assert np.abs(mhe.Ztilde[ok][:, :nx]).max() <= 10 + 1e-6   # arrival state bounds
        assert np.abs(x[ok]).max() <= 10 + 1e-6     # x̂(k) is the last row block of X̂
        assert np.array_equal(mhe.Ztilde[0], mhe.Ztilde[1]) and np.array_equal(x[0], x[1])
        for i, o in os_.items():
            xo = o.preparestate(y[i])
            assert mhe.status[i] == o.last_qp["status"], (k, i, mhe.status[i], o.last_qp["status"])
            tol = 5e-6 if mhe.iters[i] > 0 else 1e-9
            e = np.abs(x[i] - xo).max() / (1 + np.abs(xo).max())
            assert e < tol, (k, i, e, mhe.iters[i])
            worst = max(worst, e)
            nact_sample += int(mhe.iters[i] > 0)
            o.updatestate(u[i], y[i])
        nact += int((mhe.iters > 0).sum())
        mhe.updatestate(u)
    assert nact_sample >= 4, nact_sample             # the spot check exercised the IPM path
    assert nact > N                                  # the bounds really were active
    print("C3 MHE full size worst vs oracle", worst, "active solves", nact, "iteration-limit exits", stat1, "of", N * steps, "sample", sample, "active solves in sample", nact_sample)
