#!/usr/bin/env python
"""bench.py -- MPC steps/sec of the batched LinMPC step (BASELINE.json metric) on N B200s.

Headline line = BASELINE.json configs[1] (C1: batch 4096 LinMPC, 2-in/2-out random stable plants, Hp=20, Hc=5, hard u
box + soft ymax).  A "step" = one control period (`moveinput!`) of the whole batch.  The JSON line also carries, under
"configs", short timed runs of the other configurations the baseline names, each with value / e2e / roofline /
cpu_baseline: C2 (65 536 x 4-in/4-out Hp=30 Hc=10, SHARDED 65 536 / N under --gpus N: strong scaling), C4 (16 384 x
8-in/8-out Hp=50 Hc=20, hard u + hard du + soft ymin/ymax, sharded 16 384 / N) and C3 (8192 linear
MovingHorizonEstimator He=15, bounds on x, w, v, sharded 8192 / N).  Inputs of every period come from a closed-loop
trajectory recorded once, untimed (plant = model, batched SteadyKalmanFilter, setpoint steps every 25 periods;
MHE: PRBS inputs, process and measurement noise) -- "synthetic".  Legs:
  value  : inputs already in HBM, CUDA events around each launch on the launching stream, L2 flushed (256 MiB memset)
           before every timed launch;
  e2e    : the same periods through the C ABI with pinned HOST buffers (H2D + kernel + D2H per call);
  cpu_baseline : oracle/cpu_ref (restated reference path: assembly + OSQP-style ADMM) on the host cores, bounded sample;
  --impl reference : only the CPU path, with every host thread, printed in the same JSON schema.  This arm records
           its own closed-loop trajectory on the CPU (oracle/cpu_ref closed loop): it never touches libbmpc.so.
Multi-GPU (torchrun): one process per GPU, instances sharded, no data-path collective inside the solve.  The moves of
C1 are collected on every rank (north_star: "all-gather only to collect ΔŨ") by the step kernel itself: peer stores over
NVLink into symmetric memory, published with one release store of the period number per peer (epoch flags, no
cross-rank barrier per period); a reader kernel acquires the flags of period k-1 while period k runs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MPC steps/sec (batch LinMPC Hp=20 Hc=5 nu=2 ny=2)"
MHE_CFG = dict(N=8192, He=15, nx=8, nu=4, ny=4, seed=3)


def flops_model(nu, ny, nx, Hp, Hc, neps, m_ref, mean_iters):
    """Algorithmic FLOPs of one moveinput! (multiply-add = 2)."""
    nY, n = ny * Hp, nu * Hc + neps
    f_asm = 2 * nY * (nx + nu) + 2 * nY * n + 4 * nY + 2 * nx * (nx + nu) + 2 * m_ref
    f_it = m_ref * n * (n + 1) + n ** 3 / 3 + 4 * n * n + 8 * m_ref * n
    chol_it = n ** 3 / 3 + 4 * n * n
    return f_asm + mean_iters * f_it, mean_iters * chol_it, f_asm, f_it


def flops_model_mhe(nx, nym, He, m_ref, mean_iters):
    """SURVEY.md section 8d, MHE: Hessian rebuild + fresh Cholesky + unconstrained solve, then the IPM iterations."""
    n, nV = nx * (1 + He), nym * He
    f_asm = 2 * (nV + nx) * n * (n + 1) / 2 + n ** 3 / 3 + 4 * n * n + 2 * nV * n
    f_it = m_ref * n * (n + 1) + n ** 3 / 3 + 4 * n * n + 8 * m_ref * n
    return f_asm + mean_iters * f_it, f_asm, f_it


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.rows, self.stop, self.idx = [], threading.Event(), gpu_index
        self.q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                  "clocks_event_reasons.sw_power_cap")

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# =====================================================================================================
# reference arm: the reference's CPU path (restated: oracle/cpu_ref) -- no product code on this path
# =====================================================================================================
def reference_linmpc(name, n_inst, periods_rec, periods_timed, threads):
    """Closed loop of the first ``n_inst`` controllers of config ``name`` on the CPU (recording), then the timed replay
    of the last ``periods_timed`` recorded periods (fresh solver workspaces, warm-started by the periods before)."""
    from oracle import cpu_ref
    from oracle import workloads as ow
    N0, nx, nu, ny, Hp, Hc, seed = ow.CONFIGS[name]
    A, Bu, C, rng = ow.random_plants(N0, nx, nu, ny, seed)  # (= the plants of rank 0 of the GPU arm)
    ry = ow.setpoints(rng, N0, ny, periods_rec, period=25)[:, :n_inst]
    t0 = time.time()
    mpcs = ow.controllers(name, A, Bu, C, range(n_inst))
    rec = cpu_ref.closed_loop(mpcs, ry, threads=threads)
    t_prep = time.time() - t0
    mpcs = ow.controllers(name, A, Bu, C, range(n_inst))  # fresh controllers for the replay
    # the whole recording is replayed; the periods before the last ``periods_timed`` run UNTIMED and leave the ADMM
    # workspaces warm (primal / dual iterates, adapted rho), as they are in the reference's closed loop
    out = cpu_ref.run(mpcs, rec["xhat0"], rec["lastu0"], ry, threads=threads, warm=periods_rec - periods_timed)
    return dict(value=n_inst * periods_timed / out["seconds"], seconds=out["seconds"], threads=out["threads"],
                admm_iters_per_solve=float(out["iters"][periods_rec - periods_timed:].mean()), instances=n_inst,
                periods=periods_timed, prep_s=t_prep, non_converged=int((out["status"][periods_rec - periods_timed:] != 0).sum()))


def oracle_mhe_windows(n_inst, periods_moving, seed_off=0):
    """Oracle MHE loop (numpy, exact QP) for the first ``n_inst`` estimators of C3: returns the estimator objects and
    the QP of every MOVING-window period (build_qp dicts + warm start) for the CPU baseline to re-solve."""
    from oracle.mhe import MovingHorizonEstimator
    from oracle.linmpc import LinModel
    from oracle import workloads as ow
    c = MHE_CFG
    He, nx, nu, ny = c["He"], c["nx"], c["nu"], c["ny"]
    A, Bu, C, rng = ow.random_plants(max(n_inst, 1), nx, nu, ny, c["seed"] + seed_off)
    mhes, plants = [], []
    for i in range(n_inst):
        m = MovingHorizonEstimator(LinModel(A[i], Bu[i], C[i]), He=He, nint_ym=[0] * ny)
        m.setconstraint(xhatmin=[-10] * nx, xhatmax=[10] * nx, whatmin=[-0.5] * nx, whatmax=[0.5] * nx,
                        vhatmin=[-3] * ny, vhatmax=[3] * ny)
        mhes.append(m)
        plants.append(LinModel(A[i], Bu[i], C[i]))
    u = rng.choice([-1.0, 1.0], (n_inst, nu))
    windows = []
    for k in range(He + periods_moving):
        if k % 5 == 0:
            u = rng.choice([-1.0, 1.0], (n_inst, nu))
        ws = []
        for i, (m, p) in enumerate(zip(mhes, plants)):
            p.x0 = p.x0 + rng.standard_normal(nx) / nx
            y = p.evaloutput() + rng.standard_normal(ny)
            Zprev = m.Ztilde.copy()
            m.preparestate(y)
            if k >= He:
                P = m.build_qp()
                Zs = np.zeros(m.nZ)
                Zs[:nx] = m.xhat0arr_old
                Zs[nx:nx * He] = Zprev[2 * nx:nx * (He + 1)]
                P["Zs"] = Zs
                ws.append(P)
            m.updatestate(u[i], y)
            p.updatestate(u[i])
        if k >= He:
            windows.append(ws)
    return mhes, windows


def reference_mhe(n_inst, periods, threads):
    from oracle import cpu_ref
    t0 = time.time()
    mhes, windows = oracle_mhe_windows(n_inst, periods)
    t_prep = time.time() - t0
    out = cpu_ref.mhe_run(mhes, windows, threads=threads)
    return dict(value=n_inst * periods / out["seconds"], seconds=out["seconds"], threads=out["threads"],
                admm_iters_per_solve=float(out["iters"].mean()), instances=n_inst, periods=periods, prep_s=t_prep,
                non_converged=int((out["status"] == 1).sum()), infeasible=int((out["status"] == 3).sum()))


def cpu_baseline_entry(res, unit, what):
    return {"value": res["value"], "unit": unit, "cores": res["threads"], "kind": "port",
            "sample": f"{what}: first {res['instances']} instances x {res['periods']} periods ({res['seconds']:.2f} s of CPU "
                      f"wall time on {res['threads']} threads); restated reference path = per-period assembly + OSQP-style ADMM "
                      f"(OSQP binary unavailable), mean {res['admm_iters_per_solve']:.0f} ADMM iterations/solve, "
                      f"{res['non_converged']} solves at the iteration cap" +
                      (f", {res['infeasible']} windows certified primal infeasible" if "infeasible" in res else "")}


def workload_text(name):
    return {
        "C1": "BASELINE.json configs[1]: batch 4096 LinMPC, 2-in/2-out random stable LinModel (nx=4, nint_ym=[1,1] -> nxhat=6), "
              "Hp=20 Hc=5, Mwt=1 Nwt=0.1 Cwt=1e5, hard u in [-1,1] + soft ymax=0.8, setpoint steps +-1 every 25 periods, "
              "closed loop (plant = model, SteadyKalmanFilter)",
        "C2": "BASELINE.json configs[2]: batch 65536 LinMPC, 4-in/4-out random stable LinModel (nx=8 -> nxhat=12), Hp=30 Hc=10, "
              "hard u box + soft ymax, sharded over the GPUs",
        "C4": "BASELINE.json configs[4]: batch 16384 LinMPC, 8-in/8-out plant (nx=16 -> nxhat=24), Hp=50 Hc=20, hard u + hard du "
              "boxes, soft ymin/ymax, sharded over the GPUs",
        "C3": "BASELINE.json configs[3]: batch 8192 linear MovingHorizonEstimator, He=15, 4-in/4-out plant (nx=8), bounds on "
              "x (+-10), w (+-0.5), v (+-3), moving-window periods, sharded over the GPUs",
    }[name]


def run_reference(args, K, W):
    """`--impl reference`: the restated CPU path on every host thread, same metric / config / JSON schema."""
    from oracle import workloads as ow
    threads = os.cpu_count() or 1
    N, nx, nu, ny, Hp, Hc, seed = ow.CONFIGS["C1"]
    n_inst = min(args.cpu_instances, N)
    res = reference_linmpc("C1", n_inst, W + K, K, threads)
    config = {"workload": workload_text("C1"), "instances_per_gpu": N, "nu": nu, "ny": ny, "Hp": Hp, "Hc": Hc,
              "n_decision": nu * Hc + 1, "rows_reference": 2 * nu * Hp + ny * Hp + 1,
              "l2": "flushed (256 MiB memset) before every timed launch", "parallelism": "host threads"}
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": "instance-steps/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * N / res["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": cpu_baseline_entry(res, "instance-steps/s", "C1, closed loop recorded on the CPU, last K periods replayed"),
            "e2e": {"value": res["value"], "unit": "instance-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "trajectory": "recorded by the CPU closed loop of this arm (oracle/cpu_ref); no GPU library is loaded"}
    if not args.no_configs:
        cfgs = {}
        r2 = reference_linmpc("C2", 128, 8, 5, threads)
        cfgs["C2"] = {"value": r2["value"], "unit": "instance-steps/s", "cpu_baseline": cpu_baseline_entry(r2, "instance-steps/s", "C2")}
        r4 = reference_linmpc("C4", 32, 5, 3, threads)
        cfgs["C4"] = {"value": r4["value"], "unit": "instance-steps/s", "cpu_baseline": cpu_baseline_entry(r4, "instance-steps/s", "C4")}
        r3 = reference_mhe(12, 4, threads)
        cfgs["C3"] = {"value": r3["value"], "unit": "estimates/s", "cpu_baseline": cpu_baseline_entry(r3, "estimates/s", "C3 (MHE)")}
        line["configs"] = cfgs
    print(json.dumps(line))


# =====================================================================================================
# our arm
# =====================================================================================================
def shard(n_total, world, rank):
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


def build_linmpc(name, rank, world, total_periods, device):
    """Batch of this rank for config ``name`` + recorded closed-loop trajectory (untimed)."""
    import mpc_b200
    from mpc_b200 import workloads
    N0, nx, nu, ny, Hp, Hc, seed = workloads.CONFIGS[name]
    if name == "C1":
        # weak scaling: every rank runs the SAME 4096 controllers (configs[1] replicated per GPU) -- equal work per GPU by
        # construction; BMPC_RANK_SEEDS=1 gives every rank its own plants instead (the unluckiest set then paces the run)
        rs = rank if os.environ.get("BMPC_RANK_SEEDS", "0") != "0" else 0
        model, rng = workloads.random_plants(N0, nx, nu, ny, seed + 1000 * rs)
        N = N0
        ry = workloads.setpoints(rng, N, ny, total_periods, period=25)
    else:             # strong scaling: the config's batch is split over the ranks
        full, rng = workloads.random_plants(N0, nx, nu, ny, seed)
        lo, hi = shard(N0, world, rank)
        N = hi - lo
        model = mpc_b200.LinModel(full.A[lo:hi], full.Bu[lo:hi], full.C[lo:hi], N=N)
        ry = workloads.setpoints(rng, N0, ny, total_periods, period=25)[:, lo:hi]
    mpc = mpc_b200.LinMPC(model, Hp=Hp, Hc=Hc, Cwt=1e5, device=device)
    mpc.setconstraint(**workloads.constraint_kwargs(name, nu, ny))
    plant = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
    rec = dict(xhat0=[], lastu0=[], ry=[], Zin=[], iters=[], status=[], u=[])
    for k in range(total_periods):
        y = plant.evaloutput()
        mpc.preparestate(y)
        rec["xhat0"].append(mpc.estim.xhat0.copy())
        rec["lastu0"].append(mpc.batch.lastu0.copy())
        rec["Zin"].append(mpc.batch.Ztilde.copy())
        rec["ry"].append(np.ascontiguousarray(ry[k]))
        u = mpc.moveinput(ry[k])
        rec["u"].append(np.array(u, copy=True))
        rec["iters"].append(mpc.batch.iters.copy())
        rec["status"].append(mpc.batch.status.copy())
        plant.updatestate(u)
        mpc.updatestate(u, y)
    rec = {k: np.ascontiguousarray(np.stack(v)) for k, v in rec.items()}
    return mpc, model, rec


class Gather:
    """Fused all-gather of Z̃, pull protocol (bmpc_set_gather_pull) over torch symmetric memory; NCCL as the fallback.
    The step kernel stores Z̃ into its own slot buffer and publishes the period number; a one-sided pull kernel on a SIDE
    stream copies every peer's rows over NVLink while the next period computes."""

    def __init__(self, b, world, rank, N, n, dev, dist):
        import torch
        self.b, self.world, self.dist, self.mode, self.ok = b, world, dist, "none", None
        self.fused = False
        if world <= 1:
            return
        self.gather_nccl = torch.zeros((world, N, n), dtype=torch.float64, device=dev)
        self.mode = "ncclAllGather of Ztilde after every step"
        if os.environ.get("BMPC_FUSED_GATHER", "1") == "0":
            return
        if os.environ.get("BMPC_FUSED_GATHER") == "off":  # diagnostic: independent shards, no collection at all
            self.world, self.mode = 1, "off (diagnostic)"
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem
            slots = 4
            self.buf = symm_mem.empty((slots, N, n), dtype=torch.float64, device=dev)
            self.buf.zero_()
            self.localbuf = torch.zeros((slots, N, n), dtype=torch.float64, device=dev) if os.environ.get("BMPC_PULL_LOCALBUF") else None
            self.flags = symm_mem.empty((16,), dtype=torch.int64, device=dev)
            self.flags.zero_()
            hb = symm_mem.rendezvous(self.buf, dist.group.WORLD)
            hf = symm_mem.rendezvous(self.flags, dist.group.WORLD)
            self.dst = torch.zeros((world * N, n), dtype=torch.float64, device=dev)
            self.side = torch.cuda.Stream(device=dev)
            torch.cuda.synchronize()
            dist.barrier()
            b.set_gather_pull([self.localbuf.data_ptr() if (self.localbuf is not None and p == rank) else int(hb.buffer_ptrs[p]) for p in range(world)], [int(hf.buffer_ptrs[p]) for p in range(world)],
                              rank, [p * N for p in range(world + 1)], slots)
            self.fused = True
            self.mode = (f"fused, pull protocol: the step kernel's epilogue stores Ztilde into its own symmetric-memory slot buffer "
                         f"({slots} slots) and its last CTA publishes the period number in its OWN flag array (st.release.sys); a one-sided pull "
                         "kernel on a side stream polls the peers' flags and copies their rows over NVLink while the next period "
                         "runs, then acks; no cross-rank barrier, no sender-side remote stores")
        except Exception as e:  # noqa: BLE001 -- any failure of the symmetric-memory path falls back to NCCL
            sys.stderr.write(f"[bench] symmetric memory unavailable ({e!r}); using ncclAllGather\n")

    def after_step(self, z_dev):
        """Called right after a step was enqueued: collect that period's moves of all ranks."""
        if self.world <= 1:
            return
        if self.fused:
            pm = os.environ.get("BMPC_PULL_MODE", "side")  # diagnostics: "none" = publish only, "main" = pull on the step's stream
            if pm != "none":
                self.b.gather_pull(self.b.gather_epoch(), self.dst.data_ptr(), self.side.cuda_stream if pm == "side" else None)
        else:
            self.dist.all_gather_into_tensor(self.gather_nccl.view(-1), z_dev.reshape(-1))

    def join(self, stream):
        """The launching stream waits for the side stream's pulls (end of a timed region)."""
        if self.fused:
            stream.wait_stream(self.side)

    def disable(self):
        import torch
        if self.fused:
            torch.cuda.synchronize()
            self.dist.barrier()
            self.b.set_gather(None, 0)
        self.world = 1

    def verify(self, z_dev):
        """The pulled array of the LAST period against ncclAllGather of the same Z̃."""
        import torch
        if not self.fused:
            return None
        torch.cuda.synchronize()
        self.dist.all_gather_into_tensor(self.gather_nccl.view(-1), z_dev.reshape(-1))
        torch.cuda.synchronize()
        return bool(torch.equal(self.gather_nccl.view(-1), self.dst.reshape(-1))) and self.b.gather_timed_out() == 0


def time_linmpc(mpc, rec, W, K, dev, stream, flush, gather=None, world=1, dist=None, busy_passes=0, sampler=None):
    """`value` leg: device-resident inputs, CUDA events around each launch, L2 flushed before every timed launch."""
    import torch
    b = mpc.batch
    N, nu = b.N, b.nu
    tX = torch.from_numpy(rec["xhat0"]).to(dev)
    tLU = torch.from_numpy(rec["lastu0"]).to(dev)
    tRY = torch.from_numpy(rec["ry"]).to(dev)
    tZ = torch.from_numpy(rec["Zin"]).to(dev)
    tU = torch.zeros((N, nu), dtype=torch.float64, device=dev)
    tJ = torch.zeros((N,), dtype=torch.float64, device=dev)
    tS = torch.zeros((N,), dtype=torch.int32, device=dev)
    tI = torch.zeros((N,), dtype=torch.int32, device=dev)

    def launch(k):
        b.step_device(dict(xhat0=tX[k].data_ptr(), lastu0=tLU[k].data_ptr(), ry=tRY[k].data_ptr(),
                           Ztilde=tZ[k].data_ptr(), u=tU.data_ptr(), J=tJ.data_ptr(), status=tS.data_ptr(),
                           iters=tI.data_ptr()))
        if gather is not None:
            gather.after_step(tZ[k])

    tLU0, tZ0 = tLU.clone(), tZ.clone()
    for k in range(W):
        flush.zero_()
        launch(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l_before = b.launch_count()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for j in range(K):
        flush.zero_()
        ev[j][0].record(stream)
        launch(W + j)
        ev[j][1].record(stream)
    # the pulls of the last periods may still be in flight on the side stream: their completion is part of the timed work
    ev_tail = torch.cuda.Event(enable_timing=True)
    if gather is not None:
        gather.join(stream)
    ev_tail.record(stream)
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    launches = b.launch_count() - l_before
    tail_ms = ev[K - 1][1].elapsed_time(ev_tail)
    check = gather.verify(tZ[W + K - 1]) if gather is not None else None
    if gather is not None:
        gather.disable()  # the passes below (and the e2e leg) run without readers
    # extra passes without flushes keep the GPU busy long enough for the 10 Hz clock sampler (never used for `value`)
    for _ in range(busy_passes):
        tLU.copy_(tLU0)
        tZ.copy_(tZ0)
        for k in range(W, W + K):
            b.step_device(dict(xhat0=tX[k].data_ptr(), lastu0=tLU[k].data_ptr(), ry=tRY[k].data_ptr(),
                               Ztilde=tZ[k].data_ptr(), u=tU.data_ptr(), J=tJ.data_ptr(), status=tS.data_ptr(),
                               iters=tI.data_ptr()))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(c) for a, c in ev]
    dev_ms = sum(ms) + tail_ms
    if os.environ.get("BMPC_PULL_MODE"):
        sys.stderr.write(f"[bench] tail_ms {tail_ms:.4f} sum {sum(ms):.4f} min {min(ms):.4f} max {max(ms):.4f}\n")
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    return dict(dev_ms=dev_ms, launches=int(launches), wall_s=t_wall1 - t_wall0, gather_check=check)


def time_linmpc_e2e(mpc, rec, W, K, dev, world=1, dist=None):
    """`e2e` leg: the C-ABI call with pinned host buffers; u0(k-1) and Z̃ are state of the handle (io.resident), x̂0 / ry go
    up and u / status come back every period (zero-copy: the kernel reads / writes the pinned buffers over PCIe)."""
    import ctypes as C
    import torch
    from mpc_b200 import _lib
    b = mpc.batch
    N, nu, ny = b.N, b.nu, b.ny
    pin = lambda a: torch.from_numpy(a.copy()).pin_memory().numpy()
    hX, hLU, hRY, hZ = pin(rec["xhat0"]), pin(rec["lastu0"]), pin(rec["ry"]), pin(rec["Zin"])
    hU = torch.zeros((N, nu), dtype=torch.float64).pin_memory().numpy()
    hJ = torch.zeros((N,), dtype=torch.float64).pin_memory().numpy()
    hS = torch.zeros((N,), dtype=torch.int32).pin_memory().numpy()
    hI = torch.zeros((N,), dtype=torch.int32).pin_memory().numpy()
    zero_copy = int(os.environ.get("BMPC_E2E_ZEROCOPY", "1") != "0")

    def host_step(k, resident=1):
        if resident:
            io = _lib.StepIO(xhat0=hX[k].ctypes.data, ry=hRY[k].ctypes.data, u=hU.ctypes.data, status=hS.ctypes.data,
                             device_ptrs=0, sync=1, resident=1, host_mapped=zero_copy)
        else:
            io = _lib.StepIO(xhat0=hX[k].ctypes.data, lastu0=hLU[k].ctypes.data, ry=hRY[k].ctypes.data,
                             Ztilde=hZ[k].ctypes.data, u=hU.ctypes.data, J=hJ.ctypes.data, status=hS.ctypes.data,
                             iters=hI.ctypes.data, device_ptrs=0, sync=1)
        _lib.check(_lib.lib().bmpc_step(b._h, C.byref(io)))

    host_step(0, resident=0)  # loads the recorded u0(-1) / Z̃ of period 0 into the handle
    for k in range(1, W):
        host_step(k)
    # The timed region is the C-ABI call itself, as a compiled host (the Julia ccall of INTEGRATION.md) would issue it: the
    # argument structs are filled in beforehand (one per period: the host buffers differ), so that the interpreter's struct
    # construction (several microseconds in CPython) is not billed to the library.
    ios = [_lib.StepIO(xhat0=hX[k].ctypes.data, ry=hRY[k].ctypes.data, u=hU.ctypes.data, status=hS.ctypes.data,
                       device_ptrs=0, sync=1, resident=1, host_mapped=zero_copy) for k in range(W, W + K)]
    refs = [C.byref(io) for io in ios]
    step_fn, handle = _lib.lib().bmpc_step, b._h
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for r in refs:
        rc = step_fn(handle, r)
        if rc:
            _lib.check(rc)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
    du = np.abs(hU - rec["u"][W + K - 1]).max(axis=1)
    return dict(seconds=e2e_s, h2d=N * 8 * (b.nxhat + ny), d2h=N * (8 * nu + 4), zero_copy=zero_copy,
                check={"median_abs_du": float(np.median(du)), "frac_within_1e-6": float((du < 1e-6).mean())})


def time_setmodel(mpc, rec, W, K):
    """batched `setmodel!` EVERY period (adaptive / successive-linearisation MPC, reference
    src/controller/execute.jl:621-790) followed by `moveinput!`: the augmented models go up, prediction matrices + Hessian
    + its Cholesky factor + the warp kernel's row matrix are rebuilt on the device, constraints are re-pushed, then the
    step runs -- Hessian assembly literally on the per-period path.  Host buffers, wall clock."""
    import torch
    b = mpc.batch
    ms = []
    for k in range(W + K):
        b.lastu0[:] = rec["lastu0"][k]
        b.Ztilde[:] = rec["Zin"][k]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mpc._push_model()
        mpc._push()
        t1 = time.perf_counter()
        b.step(rec["xhat0"][k], ry=rec["ry"][k])
        t2 = time.perf_counter()
        if k >= W:
            ms.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    ms = np.array(ms)
    du = np.abs(b.u - rec["u"][W + K - 1]).max(axis=1)
    return {"ms_setmodel": float(np.median(ms[:, 0])), "ms_step_after_setmodel": float(np.median(ms[:, 1])),
            "value": b.N / (1e-3 * float(np.median(ms.sum(axis=1)))), "unit": "instance-(setmodel + step)s/s",
            "what": "bmpc_set_model (H2D of Â, B̂u, Ĉ; k_build_model: init_predmat + init_quadprog on the device; Cholesky of H̃) + "
                    "bmpc_set_oppoints + bmpc_set_constraints + re-layout of the warp kernel's matrices + bmpc_step, host buffers",
            "u_vs_recorded_last_period": {"median_abs_du": float(np.median(du)), "frac_within_1e-6": float((du < 1e-6).mean())}}


def measure_fp64_peak(dev):
    """cuBLAS DGEMM throughput measured in this run (MEASURED_PEAKS.json has no fp64 entry): torch.matmul fp64 4096^3."""
    import torch
    n = 4096
    a = torch.randn((n, n), dtype=torch.float64, device=dev)
    b = torch.randn((n, n), dtype=torch.float64, device=dev)
    c = torch.empty((n, n), dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def rows_reference(name, nu, ny, Hp, Hc):
    """Rows of the reference's QP for the config's setconstraint! recipe (finite bounds only, i_b) + the eps >= 0 row."""
    from mpc_b200 import workloads
    per = dict(umin=nu * Hp, umax=nu * Hp, dumin=nu * Hc, dumax=nu * Hc, ymin=ny * Hp, ymax=ny * Hp)
    return sum(per[k] for k in workloads.CONSTRAINTS[name]) + 1


def roofline_linmpc(name, b, nu, ny, Hp, Hc, mean_iters, n_total, ms_per_step, fp64_peak, hbm_peak, traffic=None, traffic_src=None):
    n = b.n
    m_ref = rows_reference(name, nu, ny, Hp, Hc)
    fl_step, fl_chol, f_asm, f_it = flops_model(nu, ny, b.nxhat, Hp, Hc, 1, m_ref, mean_iters)
    achieved = fl_step * n_total / (ms_per_step * 1e-3) / 1e12
    # per-step I/O + per-instance constants (E, K, V, B, packed H, row bounds) in the reference's layout
    alg_bytes = n_total * 8 * (b.nxhat + 2 * nu + ny + 2 * n + 2) + n_total * 8 * (b.nY * (n + b.nxhat + nu + 1) + n * (n + 1) // 2 + m_ref)
    return {"bound": "tensor", "pipe": "fp64 (DMMA m8n8k4 + DFMA; B200's FP64 tensor rate equals its FP64 vector rate)",
            "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "cuBLAS DGEMM 4096^3 (torch.matmul fp64) measured in THIS run; MEASURED_PEAKS.json has no fp64 entry",
            "rows_reference": m_ref, "algorithmic_flops_per_instance_step": fl_step,
            "cholesky_flops_per_instance_step": fl_chol,
            "achieved_cholesky_tflops": fl_chol * n_total / (ms_per_step * 1e-3) / 1e12,
            "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                    "peak_gbs": hbm_peak}}


def run_config_linmpc(name, rank, local_rank, world, dist, dev, stream, flush, fp64_peak, hbm_peak, K, W, with_cpu):
    """Short timed run of C2 / C4 (sharded): value, e2e, roofline, cpu_baseline."""
    import torch
    from mpc_b200 import workloads
    N0, nx, nu, ny, Hp, Hc, seed = workloads.CONFIGS[name]
    t0 = time.time()
    mpc, model, rec = build_linmpc(name, rank, world, W + K, local_rank)
    t_setup = time.time() - t0
    b = mpc.batch
    b.set_stream(stream.cuda_stream)
    v = time_linmpc(mpc, rec, W, K, dev, stream, flush, None, world, dist)
    e = time_linmpc_e2e(mpc, rec, W, K, dev, world, dist)
    its = rec["iters"][W:W + K]
    stats = torch.tensor([float(its.sum()), float(its.size), float((rec["status"][W:W + K] != 0).sum()), float(its.max())],
                         dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats[3:].clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        stats[3] = mx[0]
    mean_iters = float(stats[0] / stats[1])
    ms = v["dev_ms"] / K
    out = {"workload": workload_text(name), "instances_total": N0, "instances_this_rank": b.N, "n_decision": b.n,
           "value": N0 * K / (v["dev_ms"] * 1e-3), "unit": "instance-steps/s", "ms_per_step": ms, "steps": K, "warmup": W,
           "scaling": "strong" if world > 1 else "single GPU", "gpu_launches": v["launches"],
           "e2e": {"value": N0 * K / e["seconds"], "unit": "instance-steps/s", "ms_per_step": 1e3 * e["seconds"] / K,
                   "h2d_bytes_per_step": e["h2d"], "d2h_bytes_per_step": e["d2h"], "u_vs_recorded_last_period": e["check"]},
           "solver": {"mean_ipm_iters_per_instance": mean_iters, "max_ipm_iters": int(stats[3]),
                      "non_optimal_statuses": int(stats[2]), "launch": b.launch_info()},
           "roofline": roofline_linmpc(name, b, nu, ny, Hp, Hc, mean_iters, N0, ms, fp64_peak, hbm_peak),
           "setup_and_recording_s": t_setup}
    b.close()
    del mpc, rec
    torch.cuda.empty_cache()
    if with_cpu:
        ns, pr, pt = (128, 8, 5) if name == "C2" else (32, 5, 3)
        res = reference_linmpc(name, ns, pr, pt, os.cpu_count() or 1)
        out["cpu_baseline"] = cpu_baseline_entry(res, "instance-steps/s", name)
    return out


def run_config_mhe(rank, local_rank, world, dist, dev, stream, flush, fp64_peak, K, W, with_cpu):
    """C3: batch 8192 linear MHE (sharded), moving-window periods: value (device-resident y / u, CUDA events), e2e (host
    buffers through bmhe_correct / bmhe_update), roofline (SURVEY 8d MHE formula), cpu_baseline."""
    import ctypes as C
    import torch
    import mpc_b200
    from mpc_b200 import workloads, _lib
    c = MHE_CFG
    He, nx, nu, ny = c["He"], c["nx"], c["nu"], c["ny"]
    lo, hi = shard(c["N"], world, rank)
    N = hi - lo
    full, rng = workloads.random_plants(c["N"], nx, nu, ny, c["seed"])
    model = mpc_b200.LinModel(full.A[lo:hi], full.Bu[lo:hi], full.C[lo:hi], N=N)
    t0 = time.time()
    mhe = mpc_b200.MovingHorizonEstimator(model, He=He, nint_ym=[0] * ny, device=local_rank)
    mhe.setconstraint(xhatmin=[-10] * nx, xhatmax=[10] * nx, whatmin=[-0.5] * nx, whatmax=[0.5] * nx,
                      vhatmin=[-3] * ny, vhatmax=[3] * ny)
    plant = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
    rng = np.random.default_rng(c["seed"] + 100 + rank)
    # data of every period (the MHE does not feed back into the plant: the whole data set can be generated up front)
    T = He + 2 * (W + K)
    Y, U = np.zeros((T, N, ny)), np.zeros((T, N, nu))
    u = rng.choice([-1.0, 1.0], (N, nu))
    for k in range(T):
        if k % 5 == 0:
            u = rng.choice([-1.0, 1.0], (N, nu))
        plant.x0 = plant.x0 + rng.standard_normal((N, nx)) / nx
        Y[k] = plant.evaloutput() + rng.standard_normal((N, ny))
        U[k] = u
        plant.updatestate(u)
    for k in range(He):  # growing window (untimed)
        mhe.preparestate(Y[k])
        mhe.updatestate(U[k])
    t_setup = time.time() - t0
    # ---- e2e leg first (host buffers, synchronous calls): periods He .. He+W+K ----
    its, act, bad = [], [], 0
    k0 = He
    for k in range(k0, k0 + W):
        mhe.preparestate(Y[k])
        mhe.updatestate(U[k])
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    for k in range(k0 + W, k0 + W + K):
        mhe.preparestate(Y[k])
        mhe.updatestate(U[k])
        its.append(mhe.iters.copy())
        bad += int((mhe.status == 1).sum())
    e2e_s = time.perf_counter() - t1
    its = np.stack(its)
    ninf = int((mhe.status == 2).sum())
    # ---- value leg: y / u resident in HBM, caller's stream, no per-call synchronisation, CUDA events ----
    L = _lib.lib()
    _lib.check(L.bmhe_set_stream(mhe._h, C.c_void_p(stream.cuda_stream), 0))
    k1 = k0 + W + K
    tY = torch.from_numpy(np.ascontiguousarray(Y[k1:k1 + W + K])).to(dev)
    tU = torch.from_numpy(np.ascontiguousarray(U[k1:k1 + W + K])).to(dev)
    tXh = torch.zeros((N, nx), dtype=torch.float64, device=dev)
    tS = torch.zeros((N,), dtype=torch.int32, device=dev)
    tI = torch.zeros((N,), dtype=torch.int32, device=dev)
    dp = lambda t: C.cast(t.data_ptr(), _lib.c_double_p)
    ip = lambda t: C.cast(t.data_ptr(), _lib.c_int32_p)

    def period(j):
        _lib.check(L.bmhe_correct(mhe._h, dp(tY[j]), None, dp(tXh), None, None, ip(tS), ip(tI), None, None))
        _lib.check(L.bmhe_update(mhe._h, dp(tU[j])))

    lc0 = L.bmhe_launch_count(mhe._h)
    for j in range(W):
        flush.zero_()
        period(j)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    lc1 = L.bmhe_launch_count(mhe._h)
    for j in range(K):
        flush.zero_()
        ev[j][0].record(stream)
        period(W + j)
        ev[j][1].record(stream)
    torch.cuda.synchronize()
    launches = L.bmhe_launch_count(mhe._h) - lc1
    dev_ms = sum(a.elapsed_time(b2) for a, b2 in ev)
    _lib.check(L.bmhe_set_stream(mhe._h, None, 1))
    stats = torch.tensor([dev_ms, e2e_s, float(its.sum()), float(its.size), float(bad), float(ninf), float((its > 0).sum())],
                         dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats[:2].clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        stats[:2] = mx
    dev_ms, e2e_s = float(stats[0]), float(stats[1])
    mean_iters = float(stats[2] / stats[3])
    n = nx * (1 + He)
    m_ref = 2 * nx + 4 * nx * He + 2 * ny * He
    fl, f_asm, f_it = flops_model_mhe(nx, ny, He, m_ref, mean_iters)
    ms = dev_ms / K
    achieved = fl * c["N"] / (ms * 1e-3) / 1e12
    out = {"workload": workload_text("C3"), "instances_total": c["N"], "instances_this_rank": N, "n_decision": n,
           "rows_reference": m_ref, "value": c["N"] * K / (dev_ms * 1e-3), "unit": "estimates/s", "ms_per_step": ms, "steps": K,
           "warmup": W, "scaling": "strong" if world > 1 else "single GPU", "gpu_launches": int(launches),
           "e2e": {"value": c["N"] * K / e2e_s, "unit": "estimates/s", "ms_per_step": 1e3 * e2e_s / K,
                   "h2d_bytes_per_step": N * 8 * (ny + nu), "d2h_bytes_per_step": N * (8 * (nx + n + 1 + ny * He + nx * He) + 8),
                   "copies_per_step": "H2D ym, u; D2H x̂0, Z̃, J, V̂, X̂0, status, iters (preparestate! + updatestate!)"},
           "solver": {"mean_ipm_iters_per_window": mean_iters, "active_window_fraction": float(stats[6] / stats[3]),
                      "iteration_limit_exits": int(stats[4]), "infeasible_windows_last_period": int(stats[5])},
           "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA + DFMA)", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": achieved / fp64_peak, "traffic": None, "algorithmic_flops_per_estimate": fl,
                        "peak_source": "cuBLAS DGEMM 4096^3 measured in this run"},
           "setup_and_growing_window_s": t_setup}
    mhe.close()
    torch.cuda.empty_cache()
    if with_cpu:
        res = reference_mhe(12, 4, os.cpu_count() or 1)
        out["cpu_baseline"] = cpu_baseline_entry(res, "estimates/s", "C3 (MHE)")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-instances", type=int, default=4096)
    ap.add_argument("--no-configs", action="store_true", help="headline (C1) only: skip the C2 / C4 / C3 sub-runs")
    ap.add_argument("--configs", default="C2,C4,C3")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 3)
    total = W + K

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, K, W)
        return

    import torch
    from mpc_b200 import workloads
    N, nx, nu, ny, Hp, Hc, seed = workloads.CONFIGS["C1"]
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)  # a real (non-NULL) stream: kernel launches and CUDA events share it
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    fp64_peak = measure_fp64_peak(dev)
    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    hbm_peak = peaks.get("hbm_gbs")

    mpc, model, rec = build_linmpc("C1", rank, world, total, local_rank)
    b = mpc.batch
    n = b.n
    b.set_stream(stream.cuda_stream)
    gather = Gather(b, world, rank, N, n, dev, dist)
    with ClockSampler(local_rank) as clk:
        v = time_linmpc(mpc, rec, W, K, dev, stream, flush, gather, world, dist, busy_passes=60)
    # (the e2e leg below runs without the gather: a host-side caller reads u, not the peers' moves)
    e = time_linmpc_e2e(mpc, rec, W, K, dev, world, dist)
    sm = time_setmodel(mpc, rec, 2, min(K, 10)) if (world == 1 and not args.no_configs) else None
    ms_per_step = v["dev_ms"] / K
    value = world * N * K / (v["dev_ms"] * 1e-3)
    iters_rec = rec["iters"][W:W + K]
    mean_iters = float(iters_rec.mean())

    config = {"workload": workload_text("C1"), "instances_per_gpu": N, "nu": nu, "ny": ny, "Hp": Hp, "Hc": Hc,
              "n_decision": nu * Hc + 1, "rows_reference": 2 * nu * Hp + ny * Hp + 1,
              "l2": "flushed (256 MiB memset) before every timed launch",
              "parallelism": f"{world} independent shard(s) of {N} instances (weak scaling: configs[1] replicated on every GPU)"}
    traffic, traffic_src = None, None
    for f in ("ncu_r02_warp_summary.json", "ncu_r01_warp_summary.json"):
        pth = os.path.join(ROOT, "profiles", f)
        if os.path.exists(pth):
            traffic = json.load(open(pth)).get("dram_bytes_per_launch")
            traffic_src = f"static: profiles/{f} (one ncu --set full capture of this kernel; not measured in this run)"
            break
    line = {
        "metric": METRIC, "value": value, "unit": "instance-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "gather": {"mode": gather.mode, "equals_ncclAllGather": v["gather_check"]},
        "config": config, "gpu_launches": v["launches"],  # (with --gpus > 1: one step kernel + one pull kernel per period)
        "e2e": {"value": world * N * K / e["seconds"], "unit": "instance-steps/s", "h2d_bytes_per_step": e["h2d"],
                "d2h_bytes_per_step": e["d2h"], "ms_per_step": 1e3 * e["seconds"] / K,
                "copies_per_step": "H2D xhat0, ry; D2H u, status (u0(k-1), Z̃ are handle state, io.resident = 1)" +
                ("; zero-copy: the kernel reads / writes the pinned host buffers over PCIe (io.host_mapped = 1)"
                 if e["zero_copy"] else "; cudaMemcpyAsync staging copies"), "u_vs_recorded_last_period": e["check"]},
        "solver": {"mean_ipm_iters_per_instance": mean_iters, "unconstrained_exit_fraction": float((iters_rec == 0).mean()),
                   "non_optimal_statuses": int((rec["status"][W:W + K] != 0).sum()), "tol": 1e-11, "launch": b.launch_info()},
        "roofline": roofline_linmpc("C1", b, nu, ny, Hp, Hc, mean_iters, world * N, ms_per_step, fp64_peak * world, hbm_peak, traffic, traffic_src),
        "clocks": clk.summary(),
        "setmodel_every_period": sm,
        "wall_s_timed_region": v["wall_s"],
    }
    b.close()
    del mpc
    torch.cuda.empty_cache()
    if world == 1 and rank == 0:
        threads = os.cpu_count() or 1
        n_inst = min(args.cpu_instances, N)
        res = reference_linmpc("C1", n_inst, min(total, 30), min(K, 20), threads)
        line["cpu_baseline"] = cpu_baseline_entry(res, "instance-steps/s", "C1, closed loop recorded on the CPU")
        res1 = reference_linmpc("C1", 128, 25, 15, 1)
        line["cpu_baseline"]["one_core_value"] = res1["value"]
    if not args.no_configs:
        cfgs = {}
        Kc, Wc = min(K, 8), 3
        for name in [c.strip() for c in args.configs.split(",") if c.strip()]:
            try:
                if name in ("C2", "C4"):
                    cfgs[name] = run_config_linmpc(name, rank, local_rank, world, dist, dev, stream, flush, fp64_peak, hbm_peak,
                                                   Kc if name == "C2" else min(Kc, 5), Wc, with_cpu=(world == 1))
                elif name == "C3":
                    cfgs[name] = run_config_mhe(rank, local_rank, world, dist, dev, stream, flush, fp64_peak, Kc, Wc,
                                                with_cpu=(world == 1))
            except Exception as ex:  # noqa: BLE001 -- a failing sub-run must not lose the headline line
                cfgs[name] = {"error": repr(ex)}
                if world > 1:
                    raise
        line["configs"] = cfgs
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
