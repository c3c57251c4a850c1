"""Where one interior-point iteration of the general (CTA-team) kernel spends its cycles: C2 / C4 recipes at a reduced
batch, -DBMPC_PHASE_CLK build (see phase_clk.py).  usage: BMPC_LIB=.../libbmpc_clk.so python tools/studies/phase_clk_general.py C2 8192"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from mpc_b200 import workloads, _lib

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
cfg = list(workloads.CONFIGS[name]); cfg[0] = N; workloads.CONFIGS[name] = tuple(cfg)
W, K = 3, 6
mpc, model, rec = bench.build_linmpc(name, 0, 1, W + K, 0)
b = mpc.batch
L = _lib.lib()
L.bmpc_debug_phase_clk.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
out = (C.c_longlong * 32)()
L.bmpc_debug_phase_clk(b._h, out)
import time
t0 = time.perf_counter()
for k in range(W, W + K):
    b.lastu0[:] = rec["lastu0"][k]; b.Ztilde[:] = rec["Zin"][k]
    b.step(rec["xhat0"][k], ry=rec["ry"][k])
dt = (time.perf_counter() - t0) / K
_lib.check(L.bmpc_debug_phase_clk(b._h, out))
clk = np.array(list(out), dtype=np.float64)
nit, ninst = clk[16], clk[17]
names = {0: "H x + q + G'lam (residual)", 1: "reductions + convergence test + weights", 2: "build_phi (DMMA + sparse rows)", 3: "Cholesky",
         4: "G'w (rhs of a solve)", 5: "triangular solves", 6: "dense_apply + row update + step lengths", 7: "mu_aff, sigma, corrector rows",
         8: "x, s, lam update / loop edges"}
tot = clk[:9].sum()
print(f"{name} N={N}: {1e3 * dt:.2f} ms/period e2e, {nit / max(ninst, 1):.2f} iterations per IPM solve, {ninst / K:.0f} IPM solves per period, launch {b.launch_info()}")
print(f"  cycles per IPM iteration (thread 0 of the team): {tot / nit:.0f}")
for i in range(9):
    print(f"    [{i}] {names[i]:48s} {clk[i] / nit:9.0f}  {100 * clk[i] / tot:5.1f} %")
print(f"    (inside build_phi: weights + DMMA part {clk[9] / nit:9.0f} cycles; the rest = sparse rows, slack border, diagonal)")
