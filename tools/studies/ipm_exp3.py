import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import numpy as np
import ipm_exp as E
def run(opt, label, drop=True, cold=False):
    its = np.zeros((E.T, E.N), int); err = 0; bad = 0
    wsx = [None] * E.N
    for rec in E.data:
        i, k = rec["i"], rec["k"]
        G, h = rec["G"], rec["h"]
        if drop: G, h = G[:-1], h[:-1]
        x, it, lam, st = E.ipm(rec["H"], rec["q"], G, h, ws=None if cold else wsx[i], opt=opt)
        its[k, i] = it; bad += st
        err = max(err, np.abs(x - rec["x"]).max() / (1 + np.abs(rec["x"]).max()))
        wsx[i] = (x, lam) if (lam is not None and st == 0) else None
    a = its[5:]
    print("%-40s mean %.2f p90 %d p99 %d max %d  permax-mean %.1f err %.1e bad %d" % (label, a.mean(), np.percentile(a, 90), np.percentile(a, 99), a.max(), a.max(axis=1).mean(), err, bad), flush=True)
    return its
if __name__ == "__main__":
    run(None, "drop eps row, warm")
    run(None, "drop eps row, cold", cold=True)
    run(dict(adapt_tau=1.0), "drop eps row, warm, adapt tau")
    run(dict(adapt_tau=1.0), "drop eps row, cold, adapt tau", cold=True)
