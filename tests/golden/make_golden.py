"""Generates the committed golden fixtures of tests/golden/ from the pinned oracle (oracle/ is pinned to the
reference's inline known answers by tests/test_oracle_*.py; the reference itself -- Julia -- cannot run here).

    python tests/golden/make_golden.py

Fixtures (inputs AND outputs, so that neither the oracle nor the CUDA path needs the other to be checked):
  linmpc_c0_readme.npz   BASELINE configs[0]: the reference README example (README.md:38-72) -- 1 input, 2 outputs,
                         y1 = 2 e^{-20 s}/(10 s + 1), y2 = 10/(4 s + 1), Ts = 1, as a ZOH state space with a 20-sample
                         delay chain (nx = 22), Mwt = [1, 0], Nwt = [0.1], soft ymax = [Inf, 35], ry = [5, 0], 40 periods
                         of sim! (plant = model, default SteadyKalmanFilter), for Hp = 30 (package default 10 + nk) and
                         Hp = 10; Hc = 2.
  linmpc_c1_seq.npz      8 controllers of the C1 recipe (tests/helpers.py::c1_controllers(8, seed=11)), 30 closed-loop
                         periods with a setpoint switch at period 15: per-period x̂0, u0(k-1), ry -> Z̃, u, J, status.
  mhe_seq.npz            4 estimators (3 states, 2 inputs, 2 outputs, 1 measured disturbance, bounds on x̂, ŵ, v̂),
                         He = 4, 12 periods, direct = true and direct = false: y, d, u -> x̂, Z̃, J, status.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.linmpc import LinModel, LinMPC  # noqa: E402
from oracle.mhe import MovingHorizonEstimator  # noqa: E402


def readme_model():
    """ZOH realisation of the README plant: x1' = a1 x1 + b1 u (gain 2, tau 10), 20-sample delay chain on its output,
    x2' = a2 x2 + b2 u (gain 10, tau 4)."""
    a1, a2 = np.exp(-1 / 10), np.exp(-1 / 4)
    nx = 22
    A, Bu, C = np.zeros((nx, nx)), np.zeros((nx, 1)), np.zeros((2, nx))
    A[0, 0], Bu[0, 0] = a1, 1 - a1                   # x1: unit-gain lag
    A[1, 0] = 2.0                                    # delay chain input = 2 x1
    for k in range(2, 21):
        A[k, k - 1] = 1.0
    C[0, 20] = 1.0                                   # y1 = output of the 20th delay
    A[21, 21], Bu[21, 0], C[1, 21] = a2, 1 - a2, 10.0
    return LinModel(A, Bu, C, Ts=1.0)


def readme_closed_loop(Hp, steps=40):
    mpc = LinMPC(readme_model(), Hp=Hp, Hc=2, Mwt=[1, 0], Nwt=[0.1])
    mpc.setconstraint(ymax=[np.inf, 35.0])
    plant = readme_model()
    ry = np.array([5.0, 0.0])
    U, Y, X, Z, J, ST = [], [], [], [], [], []
    for _ in range(steps):
        y = plant.evaloutput()
        mpc.preparestate(y)
        X.append(mpc.estim.xhat0.copy())
        u = mpc.moveinput(ry)
        U.append(u.copy()); Y.append(y.copy()); Z.append(mpc.Ztilde.copy()); J.append(mpc.getinfo()["J"])
        ST.append(mpc.last_status)
        plant.updatestate(u)
        mpc.updatestate(u, y)
    return dict(U=np.array(U), Y=np.array(Y), xhat0=np.array(X), Z=np.array(Z), J=np.array(J), status=np.array(ST))


def c1_sequence(N=8, seed=11, steps=30):
    from helpers import c1_controllers
    mpcs, plants, rng = c1_controllers(N, seed=seed)
    r0 = rng.choice([-1.0, 1.0], (N, 2))
    rec = dict(xhat0=[], lastu0=[], ry=[], Z=[], u=[], J=[], status=[])
    for k in range(steps):
        r = r0 if k < 15 else -r0
        row = {key: [] for key in rec}
        for i, (m, p) in enumerate(zip(mpcs, plants)):
            y = p.evaloutput()
            m.preparestate(y)
            row["xhat0"].append(m.estim.xhat0.copy()); row["lastu0"].append(m.lastu0.copy()); row["ry"].append(r[i])
            u = m.moveinput(r[i])
            row["Z"].append(m.Ztilde.copy()); row["u"].append(u.copy()); row["J"].append(m.getinfo()["J"])
            row["status"].append(m.last_status)
            p.updatestate(u)
            m.updatestate(u, y)
        for key in rec:
            rec[key].append(np.array(row[key]))
    return {k: np.array(v) for k, v in rec.items()}


def mhe_models(N=4, seed=21, nx=3, nu=2, ny=2, nd=1):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((N, nx, nx))
    A *= (rng.uniform(0.5, 0.9, N) / np.abs(np.linalg.eigvals(A)).max(axis=1))[:, None, None]
    return dict(A=A, Bu=rng.standard_normal((N, nx, nu)), C=rng.standard_normal((N, ny, nx)),
                Bd=rng.standard_normal((N, nx, nd)), Dd=0.1 * rng.standard_normal((N, ny, nd)),
                uop=np.array([1.0, -2.0]), yop=np.array([5.0, 3.0]), dop=np.array([0.5])), rng


MHE_BOUNDS = dict(xhatmin=[-0.6] * 3, xhatmax=[0.6] * 3, whatmin=[-0.3] * 3, whatmax=[0.3] * 3,
                  vhatmin=[-2.5] * 2, vhatmax=[2.5] * 2)


def mhe_sequence(steps=12, He=4):
    mm, rng = mhe_models()
    N = mm["A"].shape[0]
    y = mm["yop"] + rng.standard_normal((steps, N, 2))
    d = mm["dop"] + 0.3 * rng.standard_normal((steps, N, 1))
    u = mm["uop"] + rng.standard_normal((steps, N, 2))
    out = dict(y=y, d=d, u=u, **{"model_" + k: v for k, v in mm.items()})
    for direct in (True, False):
        tag = "direct" if direct else "pred"
        es = [MovingHorizonEstimator(LinModel(mm["A"][i], mm["Bu"][i], mm["C"][i], Bd=mm["Bd"][i], Dd=mm["Dd"][i],
                                              uop=mm["uop"], yop=mm["yop"], dop=mm["dop"]), He=He, nint_ym=0,
                                     direct=direct).setconstraint(**MHE_BOUNDS) for i in range(N)]
        X, Z, J, ST = [], [], [], []
        for k in range(steps):
            xs = []
            for i, e in enumerate(es):
                xp = e.preparestate(y[k, i], d[k, i])
                xu = e.updatestate(u[k, i], y[k, i], d[k, i])
                xs.append(xp if direct else xu)      # the estimate the solved window produced
            X.append(np.array(xs)); Z.append(np.array([e.Ztilde for e in es])); J.append([e.Jval for e in es])
            ST.append([e.last_qp["status"] for e in es])
        out.update({tag + "_xhat": np.array(X), tag + "_Z": np.array(Z), tag + "_J": np.array(J),
                    tag + "_status": np.array(ST)})
    return out


if __name__ == "__main__":
    c0 = {}
    for Hp in (30, 10):
        c0.update({f"Hp{Hp}_{k}": v for k, v in readme_closed_loop(Hp).items()})
    np.savez_compressed(os.path.join(HERE, "linmpc_c0_readme.npz"), **c0)
    np.savez_compressed(os.path.join(HERE, "linmpc_c1_seq.npz"), **c1_sequence())
    np.savez_compressed(os.path.join(HERE, "mhe_seq.npz"), **mhe_sequence())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    print("C0 Hp=30: final y =", c0["Hp30_Y"][-1], " max y2 =", c0["Hp30_Y"][:, 1].max(), " statuses", set(c0["Hp30_status"]))
    print("C0 Hp=10: final y =", c0["Hp10_Y"][-1], " max y2 =", c0["Hp10_Y"][:, 1].max(), " statuses", set(c0["Hp10_status"]))
