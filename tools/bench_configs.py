"""Timing of the non-headline LinMPC configs of BASELINE.json (C2: 4x4 plant Hp=30 Hc=10; C4: 8x8 plant Hp=50 Hc=20
with hard u box, hard du box and soft ymin/ymax) through the C ABI: closed loop with plant = model, batched
SteadyKalmanFilter, setpoint steps every 25 periods.  Prints one JSON line per config.
usage: python tools/bench_configs.py C2 [N] [periods]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200
from mpc_b200 import workloads

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
N0, nx, nu, ny, Hp, Hc, seed = workloads.CONFIGS[name]
N = int(sys.argv[2]) if len(sys.argv) > 2 else N0
periods = int(sys.argv[3]) if len(sys.argv) > 3 else 30
t0 = time.time()
model, rng = workloads.random_plants(N, nx, nu, ny, seed)
mpc = mpc_b200.LinMPC(model, Hp=Hp, Hc=Hc, Cwt=1e5)
if name == "C4":
    mpc.setconstraint(umin=[-1.0] * nu, umax=[1.0] * nu, dumin=[-0.2] * nu, dumax=[0.2] * nu, ymin=[-1.2] * ny, ymax=[0.8] * ny)
else:
    mpc.setconstraint(umin=[-1.0] * nu, umax=[1.0] * nu, ymax=[0.8] * ny)
ry = workloads.setpoints(rng, N, ny, periods, period=25)
plant = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
t_setup = time.time() - t0
ms, its, bad = [], [], 0
for k in range(periods):
    y = plant.evaloutput()
    mpc.preparestate(y)
    t1 = time.perf_counter()
    u = mpc.moveinput(ry[k])
    ms.append(1e3 * (time.perf_counter() - t1))
    its.append(mpc.batch.iters.copy())
    bad += int((mpc.batch.status != 0).sum())
    plant.updatestate(u)
    mpc.updatestate(u, y)
its = np.stack(its)
skip = 3
b = mpc.batch
nrows = b.launch_info()["rows_m"]
n = b.n
f_it = nrows * n * (n + 1) + n ** 3 / 3 + 4 * n * n + 8 * nrows * n
mean_it = float(its[skip:].mean())
ms_med = float(np.median(ms[skip:]))
print(json.dumps({"config": name, "N": N, "nu": nu, "ny": ny, "nxhat": b.nxhat, "Hp": Hp, "Hc": Hc, "n": n, "rows_compiled": nrows,
                  "ms_per_period_e2e_median": ms_med, "instance_steps_per_s": N / (ms_med * 1e-3), "mean_ipm_iters": mean_it,
                  "max_ipm_iters": int(its[skip:].max()), "non_optimal": bad, "setup_s": t_setup,
                  "executed_tflops_est": (mean_it * f_it) * N / (ms_med * 1e-3) / 1e12, "launch": b.launch_info()}))
