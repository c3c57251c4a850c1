"""Timing of BASELINE.json configs[3]: batch 8192 linear MHE, He=15, 4-in/4-out plant (nx=8, nint_ym=0),
bounds x̂ in +-10, ŵ in +-0.5, v̂ in +-3; data from the plants driven by PRBS inputs and N(0, sigma) noise.
Prints one JSON line (end-to-end through the C ABI with host buffers; moving-window periods only)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200
from mpc_b200 import workloads

N = int(os.environ.get("MHE_N", 8192))
He, nx, nu, ny = 15, 8, 4, 4
steps_grow, steps_time = He, int(os.environ.get("MHE_STEPS", 20))
model, rng = workloads.random_plants(N, nx, nu, ny, seed=3)
t0 = time.time()
mhe = mpc_b200.MovingHorizonEstimator(model, He=He, nint_ym=[0] * ny)
mhe.setconstraint(xhatmin=[-10] * nx, xhatmax=[10] * nx, whatmin=[-0.5] * nx, whatmax=[0.5] * nx,
                  vhatmin=[-3] * ny, vhatmax=[3] * ny)
t_setup = time.time() - t0
plant = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
u = rng.choice([-1.0, 1.0], (N, nu))
times, iters, nact = [], [], []
for k in range(steps_grow + steps_time):
    if k % 5 == 0:
        u = rng.choice([-1.0, 1.0], (N, nu))
    plant.x0 = plant.x0 + rng.standard_normal((N, nx)) / nx  # process noise
    y = plant.evaloutput() + rng.standard_normal((N, ny))
    t1 = time.perf_counter()
    mhe.preparestate(y)
    mhe.updatestate(u)
    dt = time.perf_counter() - t1
    if k >= steps_grow:
        times.append(dt)
        iters.append(float(mhe.iters.mean()))
        nact.append(float((mhe.iters > 0).mean()))
    ninf = int((mhe.status == 2).sum())
    plant.updatestate(u)
ms = 1e3 * float(np.median(times))
err = float(np.abs(mhe.xhat0 - plant.x0).mean())
print(json.dumps({"workload": "C3: batch %d linear MHE He=15 nx=8 4x4, bounds x/w/v, moving window" % N,
                  "ms_per_period_e2e": ms, "estimates_per_s": N / (ms * 1e-3), "mean_ipm_iters": float(np.mean(iters)),
                  "active_fraction": float(np.mean(nact)), "setup_s": t_setup, "n": nx * (1 + He), "rows_m": 2 * nx + 4 * nx * He + 2 * ny * He,
                  "mean_abs_state_error": err, "infeasible_instances_last_period": ninf}))
