import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, mpc_b200
from mpc_b200 import workloads
from oracle import qp
from oracle.linmpc import LinModel as OLinModel, LinMPC as OLinMPC, ManualEstimator as OManual
N=64
_, nx, nu, ny, Hp, Hc, seed = workloads.CONFIGS["C4"]
model, rng = workloads.random_plants(2048, nx, nu, ny, seed)
xh = rng.standard_normal((2, 2048, nx + ny)) * 0.3
ry = rng.choice([-1.0, 1.0], (2, 2048, ny))
sub = mpc_b200.LinModel(model.A[:N], model.Bu[:N], model.C[:N], N=N)
c=dict(umin=[-1.0] * nu, umax=[1.0] * nu, dumin=[-0.2] * nu, dumax=[0.2] * nu, ymin=[-1.2] * ny, ymax=[0.8] * ny)
for tol in (0.0, 1e-12):
    mpc = mpc_b200.LinMPC(mpc_b200.ManualEstimator(sub), Hp=Hp, Hc=Hc, Cwt=1e5, tol=tol); mpc.setconstraint(**c)
    res=[]
    for k in range(2):
        mpc.estim.xhat0 = xh[k][:N].copy(); mpc.moveinput(ry[k][:N]); res.append((mpc.Ztilde.copy(), mpc.getinfo()))
    for i in (0,1,2,3):
        om = OLinModel(model.A[i], model.Bu[i], model.C[i]); o = OLinMPC(OManual(om), Hp=Hp, Hc=Hc, Cwt=1e5); o.setconstraint(**c)
        for k in range(2):
            o.estim.xhat0 = xh[k, i].copy(); o.moveinput(ry[k, i])
            Z, info = res[k]; oi=o.getinfo()
            ez = np.abs(Z[i] - o.Ztilde).max() / (1 + np.abs(o.Ztilde).max())
            yv=(info["Yhat"][i]-0.8-info["eps"][i]).max(); yv2=(-1.2-info["Yhat"][i]-info["eps"][i]).max()
            print("tol",tol,"inst",i,"k",k,"iters",info["iters"][i],"ez %.2e"%ez,"dJ %.3e"%(info["J"][i]-oi["J"]),"J",oi["J"],"eps gpu %.6e oracle %.6e"%(info["eps"][i],o.Ztilde[-1]),"viol %.2e %.2e"%(yv,yv2), "umax viol %.2e"%(np.abs(info["U"][i]).max()-1), "du viol %.2e"%(np.abs(info["DU"][i]).max()-0.2))
