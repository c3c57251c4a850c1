"""GPU tests of the reference-facing API: route A (on-device init_predmat/init_quadprog) against
route B (oracle matrices), the batched LinMPC mirror in closed loop against the oracle, and
size-independent properties at BASELINE.json's full C1 size."""
import numpy as np
import pytest

from oracle.linmpc import LinModel as OLinModel, LinMPC as OLinMPC, zoh_first_order
from helpers import batch_from_oracle, c1_controllers, random_plant

pytestmark = pytest.mark.gpu


def _route_a_from_oracle(mpcs, **kw):
    import mpc_b200
    m0 = mpcs[0]
    model = m0.model
    N = len(mpcs)
    b = mpc_b200.BatchLinMPC(N, model.nu, model.ny, m0.estim.nxhat, m0.Hp, m0.nb, nd=model.nd, Cwt=m0.Cwt, **kw)
    st = lambda f: np.stack([f(m) for m in mpcs])
    b.set_model(st(lambda m: m.estim.Ahat), st(lambda m: m.estim.Buhat), st(lambda m: m.estim.Chat),
                st(lambda m: m.estim.Bdhat) if model.nd else None, st(lambda m: m.estim.Ddhat) if model.nd else None,
                st(lambda m: m.estim.fophat - m.estim.xophat), st(lambda m: np.diag(m.M_Hp)),
                st(lambda m: np.diag(m.N_Hc)), st(lambda m: np.diag(m.L_Hp)))
    b.set_oppoints(st(lambda m: m.model.uop), st(lambda m: m.model.yop))
    from helpers import push_constraints
    push_constraints(b, mpcs)
    return b


@pytest.mark.parametrize("case", ["c1", "blocking_dist_terminal"])
def test_route_a_equals_route_b(case):
    rng = np.random.default_rng(5)
    mpcs = []
    for i in range(8):
        if case == "c1":
            m = random_plant(rng)
            mpc = OLinMPC(m, Hp=20, Hc=5, Cwt=1e5).setconstraint(umin=[-1, -1], umax=[1, 1], ymax=[0.8, 0.8])
        else:
            p = random_plant(rng, nx=3, nu=2, ny=2)
            m = OLinModel(p.A, p.Bu, p.C, Bd=rng.standard_normal((3, 1)), Dd=rng.standard_normal((2, 1)) * 0.1,
                          uop=[0.5, -0.2], yop=[1.0, 2.0], dop=[0.3], xop=[0.1, 0.0, -0.1], fop=[0.05, 0.1, 0.0])
            mpc = OLinMPC(m, Hp=12, Hc=[1, 2, 3], Lwt=[0.3, 0.1], Cwt=1e4)
            mpc.setconstraint(umin=[-2, -2], umax=[2, 2], dumin=[-0.7, -0.7], dumax=[0.7, 0.7], ymin=[0, 1],
                              ymax=[2.5, 3.5], xhatmin=[-5] * mpc.estim.nxhat, xhatmax=[5] * mpc.estim.nxhat,
                              c_dumin=[0.1, 0.1], c_dumax=[0.1, 0.1])
        mpcs.append(mpc)
    bA, bB = _route_a_from_oracle(mpcs), batch_from_oracle(mpcs, with_terminal=True)
    N, nd = len(mpcs), mpcs[0].model.nd
    for k in range(6):
        xh = rng.standard_normal((N, mpcs[0].estim.nxhat)) * 0.5
        ry = rng.standard_normal((N, 2)) + np.stack([m.model.yop for m in mpcs])
        d0 = rng.standard_normal((N, nd)) * 0.2 if nd else None
        lastu_before = bB.lastu0.copy()
        for b in (bA, bB):
            b.step(xh, ry=ry, d0=d0)
        assert (bA.status == 0).all() and (bB.status == 0).all()
        assert np.abs(bA.Ztilde - bB.Ztilde).max() < 1e-7 * (1 + np.abs(bB.Ztilde).max())
        assert np.abs(bA.J - bB.J).max() < 1e-9 * (1 + np.abs(bB.J).max())
        iA, iB = bA.getinfo(), bB.getinfo()
        for key in ("F", "qtilde", "r"):
            assert np.abs(iA[key] - iB[key]).max() < 1e-10 * (1 + np.abs(iB[key]).max()), key
        # ... and both against the oracle's own assembly and solution for instance 0 (initpred!, linconstraint!, optimum)
        m = mpcs[0]
        m.estim.xhat0 = xh[0].copy()
        m.lastu0 = lastu_before[0].copy()
        u_or = m.moveinput(ry[0], d=(d0[0] + m.model.dop) if nd else ())
        for key, ref in (("F", m.F), ("qtilde", m.qtilde), ("r", m.r)):
            assert np.abs(iB[key][0] - ref).max() < 1e-10 * (1 + np.abs(ref).max()), key
        assert np.abs(bB.Ztilde[0] - m.Ztilde).max() < 5e-6 * (1 + np.abs(m.Ztilde).max())
        assert np.abs(bB.u[0] - u_or).max() < 5e-6 * (1 + np.abs(u_or).max())
    assert bA.iters.max() > 0


def test_linmpc_mirror_closed_loop_vs_oracle():
    """The batched LinMPC mirror (route A + batched SKF) against N oracle controllers, 30 periods."""
    import mpc_b200
    from mpc_b200 import workloads
    N, steps = 12, 30
    model, rng = workloads.random_plants(N, 4, 2, 2, seed=11)
    mpc = mpc_b200.LinMPC(model, Hp=20, Hc=5, Cwt=1e5).setconstraint(umin=[-1, -1], umax=[1, 1], ymax=[0.8, 0.8])
    ry = workloads.setpoints(rng, N, 2, steps, period=10)
    res = mpc_b200.sim(mpc, steps, lambda k: ry[k])
    for i in range(N):
        om = OLinModel(model.A[i], model.Bu[i], model.C[i])
        o = OLinMPC(om, Hp=20, Hc=5, Cwt=1e5).setconstraint(umin=[-1, -1], umax=[1, 1], ymax=[0.8, 0.8])
        plant = OLinModel(model.A[i], model.Bu[i], model.C[i])
        for k in range(steps):
            y = plant.evaloutput()
            o.preparestate(y)
            u = o.moveinput(ry[k, i])
            assert np.abs(y - res["Y"][k, i]).max() < 2e-5, (i, k)
            assert np.abs(u - res["U"][k, i]).max() < 2e-5, (i, k)
            plant.updatestate(u)
            o.updatestate(u, y)


def test_golden_doctest_through_gpu():
    """ext/LinearMPCext.jl:252-261: u = 17.577311 (unconstrained exit of the CUDA step)."""
    import mpc_b200
    A, B, C = zoh_first_order(2, 10, 1.0, b=0.5)
    mpc = mpc_b200.LinMPC(mpc_b200.LinModel(A, B, C, Ts=1.0))
    mpc.preparestate([[1.0]])
    u = mpc.moveinput([[10.0]])
    assert round(float(u[0, 0]), 6) == 17.577311
    assert mpc.batch.iters[0] == 0 and mpc.batch.status[0] == 0


def test_full_size_c1_properties():
    """BASELINE.json configs[1] at full size (4096 instances): properties that need no oracle.
    (1) every instance solves; (2) hard input box honoured, soft output bound honoured up to eps;
    (3) results are independent of the instance order (bitwise) and of the team size (1e-9);
    (4) duplicated instances give bitwise identical results."""
    import mpc_b200
    from mpc_b200 import workloads
    N, nx, nu, ny, Hp, Hc, seed = workloads.CONFIGS["C1"]
    model, rng = workloads.random_plants(N, nx, nu, ny, seed)
    model.A[1], model.Bu[1], model.C[1] = model.A[0], model.Bu[0], model.C[0]  # duplicate instance
    xh = rng.standard_normal((N, nx + ny)) * 0.3
    xh[1] = xh[0]
    ry = rng.choice([-1.0, 1.0], (N, ny))
    ry[1] = ry[0]

    def run(perm, team=0):
        sub = mpc_b200.LinModel(model.A[perm], model.Bu[perm], model.C[perm], N=N)
        mpc = mpc_b200.LinMPC(mpc_b200.ManualEstimator(sub), Hp=Hp, Hc=Hc, Cwt=1e5, team=team)
        mpc.setconstraint(umin=[-1] * nu, umax=[1] * nu, ymax=[0.8] * ny)
        mpc.estim.xhat0 = xh[perm].copy()
        mpc.moveinput(ry[perm])
        return mpc, mpc.getinfo()

    ident = np.arange(N)
    mpc, info = run(ident)
    assert (info["status"] == 0).all()
    assert info["iters"].max() > 0 and info["iters"].max() < 50
    assert info["U"].max() <= 1 + 1e-7 and info["U"].min() >= -1 - 1e-7
    assert (info["Yhat"] - 0.8 - info["eps"][:, None]).max() <= 1e-6
    assert (info["eps"] >= -1e-12).all()
    assert np.array_equal(mpc.Ztilde[0], mpc.Ztilde[1])
    perm = np.random.default_rng(0).permutation(N)
    mpc_p, _ = run(perm)
    assert np.array_equal(mpc_p.Ztilde, mpc.Ztilde[perm])
    mpc_t, info_t = run(ident, team=32)
    assert np.abs(mpc_t.Ztilde - mpc.Ztilde).max() < 2e-6
    assert np.abs(info_t["J"] - info["J"]).max() < 1e-9 * (1 + np.abs(info["J"]).max())


def test_resident_state_equals_host_state():
    """io.resident = 1 (u0(k-1) and Z̃ are state of the handle, as mpc.lastu0 / mpc.Z̃ in linmpc.jl:3-49) gives
    bitwise the same inputs as round-tripping both through the host every call."""
    mpcs, plants, rng = c1_controllers(16, seed=21)
    bH, bR = batch_from_oracle(mpcs), batch_from_oracle(mpcs)
    for k in range(8):
        xh = rng.standard_normal((16, mpcs[0].estim.nxhat)) * 0.3
        ry = rng.choice([-1.0, 1.0], (16, 2))
        uH = bH.step(xh, ry=ry).copy()
        uR = bR.step(xh, ry=ry, resident=k > 0).copy()
        assert (bR.status == bH.status).all()
        assert np.array_equal(uH, uR), k


@pytest.mark.parametrize("team", [0, 128])
def test_fused_gather_epilogue_single_rank(team):
    """bmpc_set_gather with world = 1: the step kernel's epilogue writes Z̃ into slot (rank, instance) of the gather
    buffer (the multi-GPU run checks the same stores against ncclAllGather inside bench.py)."""
    import torch
    mpcs, plants, rng = c1_controllers(12, seed=31)
    b = batch_from_oracle(mpcs, team=team)
    gbuf = torch.zeros((1, 12, b.n), dtype=torch.float64, device="cuda:0")
    b.set_gather([gbuf.data_ptr()], 0)
    for k in range(3):
        xh = rng.standard_normal((12, mpcs[0].estim.nxhat)) * 0.3
        b.step(xh, ry=rng.choice([-1.0, 1.0], (12, 2)))
        torch.cuda.synchronize()
        assert np.array_equal(gbuf[0].cpu().numpy(), b.Ztilde)
    b.set_gather(None, 0)


@pytest.mark.parametrize("team", [0, 128])
def test_fused_estimator_equals_host_estimator(team):
    """SteadyKalmanFilter correct/predict fused into the step kernel (x̂0 owned by the handle) against the
    host-side estimator mirror driving the same controller, and against the oracle's closed loop."""
    import mpc_b200
    from mpc_b200 import workloads
    N = 16
    model, rng = workloads.random_plants(N, 4, 2, 2, seed=77)
    mk = lambda fused: mpc_b200.LinMPC(model, Hp=20, Hc=5, Cwt=1e5, team=team, fused_estimator=fused).setconstraint(
        umin=[-1, -1], umax=[1, 1], ymax=[0.8, 0.8])
    mF, mH = mk(True), mk(False)
    plantF = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
    plantH = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
    ry = workloads.setpoints(rng, N, 2, 30, period=10)
    for k in range(30):
        yF, yH = plantF.evaloutput(), plantH.evaloutput()
        mF.preparestate(yF)
        mH.preparestate(yH)
        uF, uH = mF.moveinput(ry[k]), mH.moveinput(ry[k])
        assert (mF.batch.status == 0).all() and (mH.batch.status == 0).all()
        assert np.abs(uF - uH).max() < 1e-8, (k, np.abs(uF - uH).max())
        xnext, xcorr = mF.batch.get_state()
        assert np.abs(xcorr - mH.estim.xhat0).max() < 1e-8
        plantF.updatestate(uF)
        plantH.updatestate(uH)
        mF.updatestate(uF, yF)
        mH.updatestate(uH, yH)
        assert np.abs(xnext - mH.estim.xhat0).max() < 1e-8


def test_two_handles_with_different_footprints_interleave():
    """Two handles that run the SAME kernel specialisation with different shared-memory footprints (Hp differs) step
    alternately: the per-function dynamic shared-memory limit is only ever raised."""
    rng = np.random.default_rng(8)
    groups = []
    for Hp in (24, 10):
        mpcs = []
        for _ in range(4):
            m = random_plant(rng)
            mpcs.append(OLinMPC(m, Hp=Hp, Hc=5, Cwt=1e5).setconstraint(umin=[-1, -1], umax=[1, 1], ymax=[0.8, 0.8]))
        groups.append((mpcs, batch_from_oracle(mpcs)))
    for k in range(3):
        for mpcs, b in groups:
            xh = rng.standard_normal((4, mpcs[0].estim.nxhat)) * 0.3
            b.step(xh, ry=rng.choice([-1.0, 1.0], (4, 2)))
            assert (b.status == 0).all()


def test_host_mapped_zero_copy_equals_staged_copies():
    """io.host_mapped = 1 (the kernel reads x̂0/ry from and writes u/status to the caller's page-locked host arrays over
    PCIe) gives bitwise the same result as the library's staging copies."""
    import torch
    mpcs, plants, rng = c1_controllers(32, seed=41)
    bS, bM = batch_from_oracle(mpcs), batch_from_oracle(mpcs)
    pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory().numpy()
    xh, ry = pin((32, mpcs[0].estim.nxhat), torch.float64), pin((32, 2), torch.float64)
    u, st = pin((32, 2), torch.float64), pin((32,), torch.int32)
    for k in range(6):
        xh[:] = rng.standard_normal(xh.shape) * 0.3
        ry[:] = rng.choice([-1.0, 1.0], ry.shape)
        uS = bS.step(xh, ry=ry).copy()
        bM.step_mapped(xh, ry, u, st, resident=k > 0)
        assert (st == bS.status).all() and (st == 0).all()
        assert np.array_equal(uS, u), k


@pytest.mark.parametrize("team", [0, 128])
def test_fused_kalman_filter_equals_host_and_oracle(team):
    """SURVEY 8f-1, time-varying KalmanFilter (src/estimator/kalman.jl:1235-1290): covariance recursion + gain in
    k_kf_cov around the step kernel, state update inside it, against the host-side mirror driving the same controller
    and against the oracle's KalmanFilter (x̂, P̂) on the same data.  Tolerances: u 2e-6 (two interior-point solves of problems that differ by rounding agree to the
    solver's accuracy, not bitwise), x̂ 1e-6 (the closed loop feeds those u differences back), P̂ 1e-10 (data-independent)."""
    import mpc_b200
    from mpc_b200 import workloads
    from oracle.linmpc import LinModel as OLinModel
    from oracle.mhe import KalmanFilter as OKF
    N = 12
    model, rng = workloads.random_plants(N, 4, 2, 2, seed=78)
    mk = lambda fused: mpc_b200.LinMPC(mpc_b200.KalmanFilter(model, sigmaP_0=[0.5, 0.4, 0.3, 0.2], sigmaR=[0.7, 1.3]),
                                       Hp=20, Hc=5, Cwt=1e5, team=team, fused_estimator=fused).setconstraint(
        umin=[-1, -1], umax=[1, 1], ymax=[0.8, 0.8])
    mF, mH = mk(True), mk(False)
    okf = [OKF(OLinModel(model.A[i], model.Bu[i], model.C[i]), sigmaP_0=[0.5, 0.4, 0.3, 0.2], sigmaR=[0.7, 1.3])
           for i in range(N)]
    plantF = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
    plantH = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
    ry = workloads.setpoints(rng, N, 2, 24, period=8)
    noise = 0.05 * rng.standard_normal((24, N, 2))
    for k in range(24):
        yF, yH = plantF.evaloutput() + noise[k], plantH.evaloutput() + noise[k]
        mF.preparestate(yF)
        mH.preparestate(yH)
        uF, uH = mF.moveinput(ry[k]), mH.moveinput(ry[k])
        assert (mF.batch.status == 0).all() and (mH.batch.status == 0).all()
        assert np.abs(uF - uH).max() < 2e-6, (k, np.abs(uF - uH).max())
        xnext, xcorr = mF.batch.get_state()
        assert np.abs(xcorr - mH.estim.xhat0).max() < 1e-6
        for i, o in enumerate(okf):
            o.preparestate(yH[i])
            assert np.abs(xcorr[i] - o.xhat0).max() < 1e-6 * (1 + np.abs(o.xhat0).max()), (k, i)
            o.updatestate(uH[i], yH[i])
        plantF.updatestate(uF)
        plantH.updatestate(uH)
        mF.updatestate(uF, yF)
        mH.updatestate(uH, yH)
        assert np.abs(xnext - mH.estim.xhat0).max() < 1e-6
        P = mF.batch.get_cov()
        assert np.abs(P - mH.estim.Phat).max() < 1e-10 * (1 + np.abs(P).max())
        assert np.abs(P - np.stack([o.Phat for o in okf])).max() < 1e-10 * (1 + np.abs(P).max())


def test_fused_gather_pull_two_handles_uneven_shards():
    """bmpc_set_gather_pull / bmpc_gather_pull: two handles on one device play ranks 0 and 1 of a world of 2 with UNEVEN
    shards (5 and 8 controllers).  Every period each "rank" pulls all 13 rows of Z̃ of that period -- no barrier between
    the launches, the flags carry the dependency -- and the ack back-pressure lets only 2 slots suffice."""
    import torch
    from helpers import batch_from_oracle, c1_controllers
    mpcs, plants, rng = c1_controllers(13, seed=21)
    shards = [mpcs[:5], mpcs[5:]]
    bs = [batch_from_oracle(sh) for sh in shards]
    n, slots, rows, offs = bs[0].n, 2, 13, [0, 5, 13]
    bufs = [torch.zeros((slots, b.N, n), dtype=torch.float64, device="cuda") for b in bs]
    flags = [torch.zeros(4, dtype=torch.int64, device="cuda") for _ in range(2)]
    dst = [torch.zeros((rows, n), dtype=torch.float64, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    for r, b in enumerate(bs):
        b.set_gather_pull([t.data_ptr() for t in bufs], [t.data_ptr() for t in flags], r, offs, slots)
    nxh = mpcs[0].estim.nxhat
    for k in range(6):
        ry = rng.choice([-1.0, 1.0], (13, 2))
        xh = rng.standard_normal((13, nxh)) * 0.3
        Z = []
        for r, b in enumerate(bs):
            sl = slice(offs[r], offs[r + 1])
            b.step(xh[sl], ry=ry[sl])
            Z.append(b.Ztilde.copy())
        Zall = np.concatenate(Z)
        for r, b in enumerate(bs):
            assert b.gather_epoch() == k + 1
            b.gather_pull(k + 1, dst[r].data_ptr())
        torch.cuda.synchronize()
        for r, b in enumerate(bs):
            assert b.gather_timed_out() == 0
            assert np.array_equal(dst[r].cpu().numpy(), Zall), (k, r)
    assert flags[0].cpu().numpy().tolist() == [6, 0, 6, 6]   # rank 0's data epoch, (unused), acks of readers 0 and 1
    assert flags[1].cpu().numpy().tolist() == [0, 6, 6, 6]
    with pytest.raises(Exception):
        bs[0].gather_pull(99, dst[0].data_ptr())  # a period that was never published
    with pytest.raises(Exception):
        bs[0].set_gather_pull([t.data_ptr() for t in bufs], [t.data_ptr() for t in flags], 0, [0, 4, 13], slots)  # N mismatch


def test_create_destroy_does_not_leak_device_memory():
    """bmpc_destroy / bmhe_destroy release every device allocation of the handle (52 MB leaked per C1 handle)."""
    import torch
    import mpc_b200
    from mpc_b200 import workloads
    model, rng = workloads.random_plants(512, 4, 2, 2, seed=3)

    def cycle():
        mpc = mpc_b200.LinMPC(model, Hp=20, Hc=5, Cwt=1e5).setconstraint(umin=[-1, -1], umax=[1, 1], ymax=[0.8, 0.8])
        mpc.preparestate(np.zeros((512, 2)))
        mpc.moveinput(np.ones((512, 2)))
        mpc.batch.close()
        mhe = mpc_b200.MovingHorizonEstimator(model, He=4, nint_ym=[0, 0])
        mhe.preparestate(np.zeros((512, 2)))
        mhe.updatestate(np.zeros((512, 2)))
        mhe.close()

    cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(6):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 4 * 1024 * 1024, (free0, free1)


def test_batched_setmodel_successive_linearisation_vs_oracle():
    """Batched ``setmodel!`` (reference src/controller/execute.jl:621-790) between control periods: every few
    periods each instance gets a new plant model (gain / pole drift as a successive linearisation would produce) and
    new operating points, once also new weights; the controller keeps Z̃, re-expresses u0(k-1) and the bounds, and the
    device rebuilds prediction matrices and Hessian.  Closed loop against N oracle controllers doing the same."""
    import mpc_b200
    from mpc_b200 import workloads
    from oracle.linmpc import LinModel as OLinModel, LinMPC as OLinMPC
    from oracle.mhe import KalmanFilter as OKF
    N, steps = 6, 24
    model, rng = workloads.random_plants(N, 3, 2, 2, seed=23)
    op = dict(uop=[0.5, -0.5], yop=[2.0, 1.0])
    gm = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N, **op)
    g = mpc_b200.LinMPC(mpc_b200.KalmanFilter(gm), Hp=12, Hc=3, Cwt=1e5)
    g.setconstraint(umin=[-1.5, -1.5], umax=[1.5, 1.5], ymax=[3.0, 2.2])
    os_, plants = [], []
    for i in range(N):
        o = OLinMPC(OKF(OLinModel(model.A[i], model.Bu[i], model.C[i], **op)), Hp=12, Hc=3, Cwt=1e5)
        o.setconstraint(umin=[-1.5, -1.5], umax=[1.5, 1.5], ymax=[3.0, 2.2])
        os_.append(o)
        plants.append(OLinModel(model.A[i], model.Bu[i], model.C[i], **op))
    ry = workloads.setpoints(rng, N, 2, steps, period=8) + np.array([2.0, 1.0])
    A, Bu, C = model.A.copy(), model.Bu.copy(), model.C.copy()
    worst, nact = 0.0, 0
    for k in range(steps):
        if k in (5, 11, 17):
            # new linearisation: drifted matrices, shifted operating points (and new weights once)
            A = A * rng.uniform(0.9, 1.05, (N, 1, 1))
            Bu = Bu * rng.uniform(0.8, 1.25, (N, 1, 1))
            op = dict(uop=[0.5 + 0.1 * k, -0.5], yop=[2.0 - 0.05 * k, 1.0 + 0.02 * k])
            wkw = dict(Mwt=[2.0, 1.0], Nwt=[0.2, 0.05], Lwt=[0.01, 0.0]) if k == 11 else {}
            g.setmodel(mpc_b200.LinModel(A, Bu, C, N=N, **op), **wkw)
            for i, o in enumerate(os_):
                o.setmodel(OLinModel(A[i], Bu[i], C[i], **op), **wkw)
        y = np.stack([p.evaloutput() for p in plants])
        g.preparestate(y)
        ug = g.moveinput(ry[k])
        assert (g.batch.status == 0).all(), (k, g.batch.status)
        nact += int((g.batch.iters > 0).sum())
        for i, o in enumerate(os_):
            o.preparestate(y[i])
            uo = o.moveinput(ry[k, i])
            e = np.abs(g.Ztilde[i] - o.Ztilde).max() / (1 + np.abs(o.Ztilde).max())
            eu = np.abs(ug[i] - uo).max() / (1 + np.abs(uo).max())
            assert e < 5e-6 and eu < 5e-6, (k, i, e, eu, g.batch.iters[i])
            worst = max(worst, e)
            o.updatestate(uo, y[i])
            plants[i].updatestate(uo)
        g.updatestate(ug, y)
    assert nact > 10
    print("batched setmodel: worst", worst, "active solves", nact)


def test_internalmodel_stochastic_predictions_and_kkt_output():
    """InternalModel estimator: the stochastic predictions Ŷs = Ks x̂s + Ps ŷs enter F through io.Yhat_s
    (predictstoch!, src/controller/execute.jl:321-327).  The reference's known answer (test/3:159-176: constant output
    disturbance, u -> 2, ym -> 15) through the CUDA path, and a constrained batch against the oracle; io.kkt reports the
    relative KKT residuals of every returned iterate."""
    import mpc_b200
    from mpc_b200 import workloads
    from oracle.linmpc import InternalModel as OIM, LinModel as OLinModel, LinMPC as OLinMPC, zoh_first_order
    A, B, C = zoh_first_order(5, 2, 3.0)
    gm = mpc_b200.LinModel(A, B, C, Ts=3.0, yop=[10])
    g = mpc_b200.LinMPC(mpc_b200.InternalModel(gm))
    plant = OLinModel(A, B, C, Ts=3.0, yop=[10])
    for i in range(25):
        ym = plant.evaloutput() - 5
        g.preparestate([ym])
        u = g.moveinput([[15]])
        g.updatestate(u, [ym])
        plant.updatestate(u[0])
    assert abs(u[0, 0] - 2) < 1e-2 and abs(ym[0] - 15) < 1e-2
    # constrained batch vs the oracle
    N, steps = 6, 14
    model, rng = workloads.random_plants(N, 3, 2, 2, seed=31)
    kw = dict(Hp=12, Hc=3, Cwt=1e5)
    cons = dict(umin=[-1, -1], umax=[1, 1], ymax=[0.6, 0.6])
    g = mpc_b200.LinMPC(mpc_b200.InternalModel(model), **kw).setconstraint(**cons)
    os_ = [OLinMPC(OIM(OLinModel(model.A[i], model.Bu[i], model.C[i])), **kw).setconstraint(**cons) for i in range(N)]
    plants = [OLinModel(model.A[i], model.Bu[i], model.C[i]) for i in range(N)]
    ry = workloads.setpoints(rng, N, 2, steps, period=7)
    dist = rng.uniform(-0.3, 0.3, (N, 2))   # constant output disturbances
    nact = 0
    for k in range(steps):
        y = np.stack([p.evaloutput() for p in plants]) + dist
        g.preparestate(y)
        ug = g.moveinput(ry[k])
        b = g.batch
        assert (b.status == 0).all()
        assert (b.kkt >= 0).all() and (b.kkt[b.iters == 0] == 0).all() and b.kkt.max() < 1e-7, b.kkt.max()
        nact += int((b.iters > 0).sum())
        for i, o in enumerate(os_):
            o.preparestate(y[i])
            uo = o.moveinput(ry[k, i])
            assert np.abs(g.Ztilde[i] - o.Ztilde).max() < 5e-6 * (1 + np.abs(o.Ztilde).max()), (k, i)
            o.updatestate(uo, y[i])
            plants[i].updatestate(uo)
        g.updatestate(ug, y)
    assert nact > 10
