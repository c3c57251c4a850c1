"""The two generators of the synthetic workloads (product side: modelpredictivecontrol.jl_b200/workloads.py; oracle
side: oracle/workloads.py, used by the CPU baseline and by `bench.py --impl reference`) produce the same plants,
setpoints and constraint recipes; the CPU closed loop (oracle/cpu_ref, no GPU code) tracks the exact oracle loop."""
import numpy as np

from oracle import cpu_ref
from oracle import workloads as ow


def test_generators_agree():
    import mpc_b200
    from mpc_b200 import workloads as pw
    assert pw.CONFIGS == ow.CONFIGS and pw.CONSTRAINTS == ow.CONSTRAINTS
    for name in ("C1", "C2", "C4"):
        _, nx, nu, ny, Hp, Hc, seed = ow.CONFIGS[name]
        N = 16
        model, rng_p = pw.random_plants(N, nx, nu, ny, seed)
        A, Bu, C, rng_o = ow.random_plants(N, nx, nu, ny, seed)
        assert np.array_equal(model.A, A) and np.array_equal(model.Bu, Bu) and np.array_equal(model.C, C)
        assert np.array_equal(pw.setpoints(rng_p, N, ny, 60), ow.setpoints(rng_o, N, ny, 60))
        assert pw.constraint_kwargs(name, nu, ny) == ow.constraint_kwargs(name, nu, ny)


def test_cpu_closed_loop_tracks_exact_oracle():
    name, N, steps = "C1", 6, 40
    _, nx, nu, ny, Hp, Hc, seed = ow.CONFIGS[name]
    A, Bu, C, rng = ow.random_plants(N, nx, nu, ny, seed)
    ry = ow.setpoints(rng, N, ny, steps, period=15)
    rec = cpu_ref.closed_loop(ow.controllers(name, A, Bu, C, range(N)), ry, threads=2)
    # exact oracle loop on the same plants
    mpcs = ow.controllers(name, A, Bu, C, range(N))
    from oracle.linmpc import LinModel
    plants = [LinModel(A[i], Bu[i], C[i]) for i in range(N)]
    U = np.zeros((steps, N, nu))
    X = np.zeros((steps, N, mpcs[0].estim.nxhat))
    for k in range(steps):
        for i, (m, p) in enumerate(zip(mpcs, plants)):
            y = p.evaloutput()
            m.preparestate(y)
            X[k, i] = m.estim.xhat0
            u = m.moveinput(ry[k, i])
            U[k, i] = u
            m.updatestate(u, y)
            p.updatestate(u)
    assert np.array_equal(rec["lastu0"][1:], rec["u"][:-1])
    # the ADMM loop (eps 1e-3) follows the exact loop at the solver's accuracy
    assert np.abs(rec["u"] - U).mean() < 2e-2 and np.abs(rec["xhat0"] - X).mean() < 2e-2
    # replaying the recorded inputs reproduces the loop's own moves (fresh solver workspaces, same iterates)
    rep = cpu_ref.run(ow.controllers(name, A, Bu, C, range(N)), rec["xhat0"], rec["lastu0"], ry, threads=2)
    assert np.abs(rep["u"] - rec["u"]).max() < 1e-9
