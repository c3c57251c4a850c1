"""Kernel time vs batch size N (C1 recipe): separates the tail (max-iteration instance) from throughput."""
import sys, os, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import torch
import bench
from mpc_b200 import workloads
for N in [int(a) for a in sys.argv[1:]] or [512, 2368, 4096, 8192, 16384]:
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
    for k in range(W):
        flush.zero_(); launch(k)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for j in range(K):
        flush.zero_(); ev[j][0].record(stream); launch(W + j); ev[j][1].record(stream)
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(c) for a, c in ev])
    it = rec["iters"][W:W + K]
    print("N %6d  ms/step mean %.4f min %.4f max %.4f | iters mean %.2f  per-period max mean %.1f | %.2f M steps/s | launch %s" % (
        N, ms.mean(), ms.min(), ms.max(), it.mean(), it.max(axis=1).mean(), N / ms.mean() / 1e3, b.launch_info()), flush=True)
    # correlation of per-step time with per-period max iterations
    mx = it.max(axis=1)
    for v in sorted(set(mx)):
        print("    max iters %2d: %.4f ms (%d periods)" % (v, ms[mx == v].mean(), (mx == v).sum()))
