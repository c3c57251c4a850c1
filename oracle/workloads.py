"""Synthetic workloads of BASELINE.json's configs as ORACLE objects (numpy only, no GPU code): the same seeded random
stable plants and constraint recipes modelpredictivecontrol.jl_b200/workloads.py generates for the CUDA path
(tests/test_workloads.py asserts the two generators agree bit for bit).  Used by bench.py's cpu_baseline leg and by
`bench.py --impl reference`, which must not touch the product library.  TEST / BENCH INFRASTRUCTURE ONLY."""
import numpy as np

from .linmpc import LinModel, LinMPC

CONFIGS = {
    # name: (N, nx, nu, ny, Hp, Hc, seed)
    "C1": (4096, 4, 2, 2, 20, 5, 1),
    "C2": (65536, 8, 4, 4, 30, 10, 2),
    "C4": (16384, 16, 8, 8, 50, 20, 4),
}
# setconstraint! recipe of each config (hard input boxes, soft output bounds; C4 adds hard increment bounds and ymin)
CONSTRAINTS = {
    "C1": dict(umin=-1.0, umax=1.0, ymax=0.8),
    "C2": dict(umin=-1.0, umax=1.0, ymax=0.8),
    "C4": dict(umin=-1.0, umax=1.0, dumin=-0.2, dumax=0.2, ymin=-1.2, ymax=0.8),
}


def random_plants(N, nx, nu, ny, seed, rho=(0.5, 0.95)):
    """A ~ N(0,1) scaled to spectral radius U(0.5, 0.95); Bu, C ~ N(0,1).  Returns (A, Bu, C, rng)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((N, nx, nx))
    rad = np.abs(np.linalg.eigvals(A)).max(axis=1)
    A *= (rng.uniform(rho[0], rho[1], N) / rad)[:, None, None]
    Bu = rng.standard_normal((N, nx, nu))
    C = rng.standard_normal((N, ny, nx))
    return A, Bu, C, rng


def setpoints(rng, N, ny, steps, period=25):
    """Setpoint steps ry in {-1,+1}^ny switching every ``period`` control periods: (steps, N, ny)."""
    nseg = (steps + period - 1) // period
    seg = rng.choice([-1.0, 1.0], (nseg, N, ny))
    return np.repeat(seg, period, axis=0)[:steps]


def constraint_kwargs(name, nu, ny):
    sizes = dict(umin=nu, umax=nu, dumin=nu, dumax=nu, ymin=ny, ymax=ny)
    return {k: [v] * sizes[k] for k, v in CONSTRAINTS[name].items()}


def controllers(name, A, Bu, C, idx):
    """Oracle LinMPC controllers of the instances ``idx`` of config ``name`` (Mwt=1, Nwt=0.1, Cwt=1e5, nint_ym=1)."""
    _, nx, nu, ny, Hp, Hc, _ = CONFIGS[name]
    out = []
    for i in idx:
        mpc = LinMPC(LinModel(A[i], Bu[i], C[i]), Hp=Hp, Hc=Hc, Cwt=1e5)
        mpc.setconstraint(**constraint_kwargs(name, nu, ny))
        out.append(mpc)
    return out
