"""Where one interior-point iteration of step_warp spends its cycles (C1 recipe): a -DBMPC_PHASE_CLK build of the
library accumulates, per phase, the cycles lane 0 sees between marks.  Build the instrumented library first:
  BMPC_NVCC_EXTRA=-DBMPC_PHASE_CLK BMPC_OUT=$PWD/tools/studies/build/libbmpc_clk.so python modelpredictivecontrol.jl_b200/build.py
Usage (GPU): BMPC_LIB=$PWD/tools/studies/build/libbmpc_clk.so python tools/studies/phase_clk.py [N ...]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from mpc_b200 import workloads, _lib

NAMES = {0: "row weights -> smem", 1: "H x + G'[lam, d rp]", 2: "reduction + convergence test", 3: "Phi by DMMA + row reload",
         4: "Cholesky", 5: "factor rows -> smem -> columns", 6: "solve (predictor)", 7: "G dx, step lengths, reduction (pred)",
         8: "corrector rhs G'w", 9: "solve (corrector)", 10: "G dx, step lengths, reduction (corr)", 11: "update x, s, lam (loop tail)",
         12: "stage 1 (initpred!, linconstraint!)", 13: "wait for the TMA copy", 14: "q, unconstrained exit, feasibility",
         15: "IPM start (incl. warm start)", 16: "loop exit", 17: "stage 4 (getinput!, outputs)", 23: "work-queue fetch", 21: "stage 0a: order entry -> instance index", 22: "stage 0b: TMA issue + prologue loads issued", 18: "stage 1a: prologue loads arrive (x̂0, u, Z̃, bounds, K/V columns)", 19: "stage 1b: prediction rows F, tY", 20: "stage 1c: L2 prefetch / terminal / custom rows", 12: "stage 1d: linconstraint! (h, q weights)"}

for N in [int(a) for a in sys.argv[1:]] or [4096, 296]:
    workloads.CONFIGS["C1"] = (N, 4, 2, 2, 20, 5, 1)
    W, K = 5, 40
    mpc, model, rec = bench.build_linmpc("C1", 0, 1, W + K, 0)
    b = mpc.batch
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); b.set_stream(stream.cuda_stream)
    tX = torch.from_numpy(rec["xhat0"]).to(dev); tLU = torch.from_numpy(rec["lastu0"]).to(dev)
    tRY = torch.from_numpy(rec["ry"]).to(dev); tZ = torch.from_numpy(rec["Zin"]).to(dev)
    tU = torch.zeros((N, 2), dtype=torch.float64, device=dev); tJ = torch.zeros((N,), dtype=torch.float64, device=dev)
    tS = torch.zeros((N,), dtype=torch.int32, device=dev); tI = torch.zeros((N,), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    def launch(k):
        b.step_device(dict(xhat0=tX[k].data_ptr(), lastu0=tLU[k].data_ptr(), ry=tRY[k].data_ptr(), Ztilde=tZ[k].data_ptr(),
                           u=tU.data_ptr(), J=tJ.data_ptr(), status=tS.data_ptr(), iters=tI.data_ptr()))
    L = _lib.lib()
    L.bmpc_debug_phase_clk.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    out = (C.c_longlong * 32)()
    for k in range(W):
        flush.zero_(); launch(k)
    torch.cuda.synchronize()
    L.bmpc_debug_phase_clk(b._h, out)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for j in range(K):
        flush.zero_(); ev[j][0].record(stream); launch(W + j); ev[j][1].record(stream)
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(c) for a, c in ev]).mean()
    _lib.check(L.bmpc_debug_phase_clk(b._h, out))
    clk = np.array(list(out), dtype=np.float64)
    ninst, nit = clk[24], clk[25]
    print(f"N {N}: {ms:.4f} ms/step (instrumented build), {ninst:.0f} instance-steps, {nit / ninst:.2f} iterations each; launch {b.launch_info()}")
    it_tot = sum(clk[i] for i in range(12))
    print(f"  cycles per IPM iteration: {it_tot / nit:.0f}   per instance outside the loop: {(sum(clk[12:24])) / ninst:.0f}")
    for i in range(12):
        print(f"    [{i:2d}] {NAMES[i]:45s} {clk[i] / nit:8.0f} cycles/iteration  {100 * clk[i] / it_tot:5.1f} %")
    for i in (23, 21, 22, 18, 19, 20, 12, 13, 14, 15, 16, 17):
        print(f"    [{i:2d}] {NAMES[i]:45s} {clk[i] / ninst:8.0f} cycles/instance")
    b.close()
