"""C1: histogram of IPM iterations per instance-period and how the per-launch time follows the slow instances."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
W, K = 5, 100
mpc, model, rec = bench.build_linmpc("C1", 0, 1, W + K, 0)
b = mpc.batch
N = rec["iters"].shape[1]
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
its = []
for j in range(K):
    flush.zero_(); ev[j][0].record(stream); launch(W + j); ev[j][1].record(stream)
    its.append(tI.cpu().numpy().copy())
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(c) for a, c in ev])
its = np.array(its)
print("mean ms %.4f  mean iters %.2f" % (ms.mean(), its.mean()))
h = np.bincount(its.ravel(), minlength=52)
print("iteration histogram (count over %d instance-periods):" % its.size)
print(" ".join("%d:%d" % (i, c) for i, c in enumerate(h) if c))
for j in range(K):
    it = its[j]
    prev = its[j - 1] if j else it
    slow = it >= 9
    print("period %3d  %.4f ms  max %2d  mean %.2f  n>=9 %4d  n>=13 %3d  predicted(prev>=9) %4d  hit %4d  sum_iters %d" % (
        W + j, ms[j], it.max(), it.mean(), slow.sum(), (it >= 13).sum(), (prev >= 9).sum(), (slow & (prev >= 9)).sum(), it.sum()))
