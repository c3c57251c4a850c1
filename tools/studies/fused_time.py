import sys, os
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import torch, bench
W, K = 5, 60
mpc, model, rec = bench.build_linmpc("C1", 0, 1, W + K, 0)
b = mpc.batch; est = mpc.estim; N = 4096
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); b.set_stream(stream.cuda_stream)
tX = torch.from_numpy(rec["xhat0"]).to(dev); tLU = torch.from_numpy(rec["lastu0"]).to(dev); tY = torch.from_numpy(rec["y0m"]).to(dev)
tRY = torch.from_numpy(rec["ry"]).to(dev); tZ = torch.from_numpy(rec["Zin"]).to(dev)
tU = torch.zeros((N, 2), dtype=torch.float64, device=dev); tJ = torch.zeros((N,), dtype=torch.float64, device=dev)
tS = torch.zeros((N,), dtype=torch.int32, device=dev); tI = torch.zeros((N,), dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
b.set_estimator(est.Ahat, est.Buhat, est.Cmhat, est.Khat, None, None, est.fophat - est.xophat)
for mode in ("xhat0", "fused", "xhat0", "fused"):
    b.set_state(np.zeros((N, b.nxhat)))
    tLU2, tZ2 = tLU.clone(), tZ.clone()
    def launch(k):
        src = dict(xhat0=tX[k].data_ptr()) if mode == "xhat0" else dict(y0m=tY[k].data_ptr())
        b.step_device(dict(lastu0=tLU2[k].data_ptr(), ry=tRY[k].data_ptr(), Ztilde=tZ2[k].data_ptr(), u=tU.data_ptr(), J=tJ.data_ptr(), status=tS.data_ptr(), iters=tI.data_ptr(), **src))
    for k in range(W):
        flush.zero_(); launch(k)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    its = []
    for j in range(K):
        flush.zero_(); ev[j][0].record(stream); launch(W + j); ev[j][1].record(stream)
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(c) for a, c in ev])
    print(mode, "ms mean %.4f min %.4f" % (ms.mean(), ms.min()), "last iters mean", tI.double().mean().item(), "max", tI.max().item())
