import sys, os, pickle
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import mpc_b200
from mpc_b200 import workloads
name = sys.argv[1]; N = int(sys.argv[2]); periods = int(sys.argv[3])
N0, nx, nu, ny, Hp, Hc, seed = workloads.CONFIGS[name]
model, rng = workloads.random_plants(N, nx, nu, ny, seed)
mpc = mpc_b200.LinMPC(model, Hp=Hp, Hc=Hc, Cwt=1e5)
mpc.setconstraint(umin=[-1.0] * nu, umax=[1.0] * nu, ymax=[0.8] * ny)
ry = workloads.setpoints(rng, N, ny, periods, period=25)
plant = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
found = []
for k in range(periods):
    y = plant.evaloutput()
    mpc.preparestate(y)
    xh = mpc.estim.xhat0.copy(); lu = mpc.batch.lastu0.copy(); Zin = mpc.batch.Ztilde.copy()
    u = mpc.moveinput(ry[k])
    st = mpc.batch.status; it = mpc.batch.iters
    bad = np.nonzero(st != 0)[0]
    for i in bad:
        print("period", k, "inst", i, "status", st[i], "iters", it[i], "J", mpc.batch.J[i], "Z", mpc.batch.Ztilde[i][:4], flush=True)
        found.append(dict(k=k, i=int(i), A=model.A[i], Bu=model.Bu[i], C=model.C[i], xhat0=xh[i], lastu0=lu[i], ry=ry[k][i], Zin=Zin[i], Z=mpc.batch.Ztilde[i].copy(), status=int(st[i]), iters=int(it[i])))
    plant.updatestate(u); mpc.updatestate(u, y)
print("iters hist", np.bincount(it, minlength=52))
os.makedirs("/root/repo/gpurun_out", exist_ok=True)
pickle.dump(found, open("/root/repo/gpurun_out/bad_%s.pkl" % name, "wb"))
