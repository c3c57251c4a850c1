"""Generate C1 closed-loop QPs (v-space) with exact solutions for IPM-variant experiments (CPU)."""
import sys, pickle, time
import numpy as np
_R = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))); sys.path.insert(0, _R); sys.path.insert(0, _R + "/tests")
from helpers import c1_controllers
from vspace_model import build_qp_v
N, T = int(sys.argv[1]), int(sys.argv[2])
mpcs, plants, rng = c1_controllers(N, seed=11)
data = []
r = rng.choice([-1.0, 1.0], (N, 2))
t0 = time.time()
for k in range(T):
    if k % 25 == 0 and k > 0:
        r = rng.choice([-1.0, 1.0], (N, 2))
    for i, (m, p) in enumerate(zip(mpcs, plants)):
        y = p.evaloutput()
        m.preparestate(y)
        u = m.moveinput(r[i])
        H, q, G, h, Dt = build_qp_v(m)
        xv = np.linalg.solve(Dt, m.Ztilde)
        data.append(dict(i=i, k=k, H=H, q=q, G=G, h=h, x=xv))
        m.updatestate(u, y)
        p.updatestate(u)
    print(k, time.time() - t0, flush=True)
pickle.dump(data, open("/tmp/mpcdata/qps_%d_%d.pkl" % (N, T), "wb"))
