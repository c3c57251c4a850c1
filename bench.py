#!/usr/bin/env python
"""bench.py -- MPC steps/sec of the batched LinMPC step (BASELINE.json metric) on N B200s.

A "step" = one control period (`moveinput!`) of the whole batch: BASELINE.json configs[1]
(batch 4096 LinMPC, 2-in/2-out random stable plants, Hp=20, Hc=5, hard u box + soft ymax).
Inputs of every period (x̂0, u0(k-1), ry, previous Z̃) come from a closed-loop trajectory recorded
once, untimed, with the same controller (plant = model, batched SteadyKalmanFilter, setpoint steps
every 25 periods) -- "synthetic".  Legs:
  value  : kernel-resident throughput, inputs already in HBM, CUDA events around each launch on the
           launching stream, L2 flushed (256 MiB memset) before every timed launch;
  e2e    : the same periods through the C ABI with pinned HOST buffers (H2D + kernel + D2H per call);
  cpu_baseline : oracle/cpu_ref (restated reference path: assembly + OSQP-style ADMM) on the host cores;
  --impl reference : only the CPU leg, with every host thread, printed in the same JSON schema.
Multi-GPU (torchrun): one process per GPU, instances sharded (independent batches, weak scaling), the
only collective is an NCCL all-gather of Z̃ after every step (north_star), timed inside the region.
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

WORKLOAD = "C1"


def flops_model(nu, ny, nx, Hp, Hc, neps, m_ref, mean_iters):
    """Algorithmic FLOPs of one moveinput! (multiply-add = 2)."""
    nY, n = ny * Hp, nu * Hc + neps
    f_asm = 2 * nY * (nx + nu) + 2 * nY * n + 4 * nY + 2 * nx * (nx + nu) + 2 * m_ref
    f_it = m_ref * n * (n + 1) + n ** 3 / 3 + 4 * n * n + 8 * m_ref * n
    chol_it = n ** 3 / 3 + 4 * n * n
    return f_asm + mean_iters * f_it, mean_iters * chol_it, f_asm, f_it


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


def build_workload(rank, total_periods, device=0):
    """C1 batch on this rank + recorded closed-loop trajectory (untimed)."""
    import mpc_b200
    from mpc_b200 import workloads
    N, nx, nu, ny, Hp, Hc, seed = workloads.CONFIGS[WORKLOAD]
    model, rng = workloads.random_plants(N, nx, nu, ny, seed + 1000 * rank)
    mpc = mpc_b200.LinMPC(model, Hp=Hp, Hc=Hc, Cwt=1e5, device=device)
    mpc.setconstraint(umin=[-1.0] * nu, umax=[1.0] * nu, ymax=[0.8] * ny)
    ry = workloads.setpoints(rng, N, ny, total_periods, period=25)
    plant = mpc_b200.LinModel(model.A, model.Bu, model.C, N=N)
    rec = dict(xhat0=[], lastu0=[], ry=[], Zin=[], iters=[], status=[], u=[], y0m=[])
    for k in range(total_periods):
        y = plant.evaloutput()
        rec["y0m"].append(y - model.yop)
        mpc.preparestate(y)
        rec["xhat0"].append(mpc.estim.xhat0.copy())
        rec["lastu0"].append(mpc.batch.lastu0.copy())
        rec["Zin"].append(mpc.batch.Ztilde.copy())
        rec["ry"].append(ry[k].copy())
        u = mpc.moveinput(ry[k])
        rec["u"].append(np.array(u, copy=True))
        rec["iters"].append(mpc.batch.iters.copy())
        rec["status"].append(mpc.batch.status.copy())
        plant.updatestate(u)
        mpc.updatestate(u, y)
    rec = {k: np.ascontiguousarray(np.stack(v)) for k, v in rec.items()}
    return mpc, model, rec


def cpu_leg(model, rec, periods, n_inst, threads, Hp, Hc):
    """Restated reference CPU path on a bounded sample of the same workload."""
    from oracle import cpu_ref
    from oracle.linmpc import LinModel as OLinModel, LinMPC as OLinMPC
    nu, ny = model.nu, model.ny
    mpcs = []
    for i in range(n_inst):
        o = OLinMPC(OLinModel(model.A[i], model.Bu[i], model.C[i]), Hp=Hp, Hc=Hc, Cwt=1e5)
        o.setconstraint(umin=[-1.0] * nu, umax=[1.0] * nu, ymax=[0.8] * ny)
        mpcs.append(o)
    sl = slice(periods[0], periods[1])
    out = cpu_ref.run(mpcs, rec["xhat0"][sl, :n_inst], rec["lastu0"][sl, :n_inst], rec["ry"][sl, :n_inst], threads=threads)
    nsteps = periods[1] - periods[0]
    return dict(value=n_inst * nsteps / out["seconds"], seconds=out["seconds"], threads=out["threads"],
                admm_iters_per_solve=float(out["iters"].mean()), instances=n_inst, periods=nsteps,
                ms_per_period_sample=1e3 * out["seconds"] / nsteps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-instances", type=int, default=4096)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 3)
    total = W + K

    import torch
    from mpc_b200 import workloads
    N, nx, nu, ny, Hp, Hc, seed = workloads.CONFIGS[WORKLOAD]
    config = {"workload": "BASELINE.json configs[1]: batch 4096 LinMPC, 2-in/2-out random stable LinModel (nx=4, "
                          "nint_ym=[1,1] -> nxhat=6), Hp=20 Hc=5, Mwt=1 Nwt=0.1 Cwt=1e5, hard u in [-1,1] + soft ymax=0.8, "
                          "setpoint steps +-1 every 25 periods, closed loop (plant = model, SteadyKalmanFilter)",
              "instances_per_gpu": N, "nu": nu, "ny": ny, "Hp": Hp, "Hc": Hc, "n_decision": nu * Hc + 1,
              "rows_reference": 2 * nu * Hp + ny * Hp + 1, "l2": "flushed (256 MiB memset) before every timed launch",
              "parallelism": f"{world} independent shard(s) of {N} instances"}

    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's own CPU implementation cannot run here (no Julia / OSQP in the image): its restated
        # path (oracle/cpu_ref) is timed instead, with every host thread, on a bounded sample of the workload.
        import mpc_b200  # the trajectory is generated once with the GPU controller (untimed)
        mpc, model, rec = build_workload(0, total)
        threads = os.cpu_count() or 1
        n_inst = min(args.cpu_instances, N)
        t_all = []
        for rep in range(1):
            res = cpu_leg(model, rec, (W, W + K), n_inst, threads, Hp, Hc)
            t_all.append(res)
        res = t_all[0]
        ms_per_step = 1e3 * N / res["value"]  # time one full-batch period would take at this rate
        line = {"impl": "reference", "metric": "MPC steps/sec (batch LinMPC Hp=20 Hc=5 nu=2 ny=2)", "value": res["value"],
                "unit": "instance-steps/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": res["value"], "unit": "instance-steps/s", "cores": res["threads"], "kind": "port",
                                 "sample": f"first {n_inst} of {N} instances x {K} recorded periods; restated OSQP-style "
                                           f"ADMM (OSQP binary unavailable), mean {res['admm_iters_per_solve']:.1f} ADMM "
                                           "iterations/solve"},
                "e2e": {"value": res["value"], "unit": "instance-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import mpc_b200
    mpc, model, rec = build_workload(rank, total, device=local_rank)
    b = mpc.batch
    n = b.n
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)  # a real (non-NULL) stream: kernel launches and CUDA events share it
    torch.cuda.set_stream(stream)
    b.set_stream(stream.cuda_stream)
    # ---- device-resident inputs for the `value` leg ----
    tX = torch.from_numpy(rec["xhat0"]).to(dev)
    tLU = torch.from_numpy(rec["lastu0"]).to(dev)
    tRY = torch.from_numpy(rec["ry"]).to(dev)
    tZ = torch.from_numpy(rec["Zin"]).to(dev)
    tU = torch.zeros((N, nu), dtype=torch.float64, device=dev)
    tJ = torch.zeros((N,), dtype=torch.float64, device=dev)
    tS = torch.zeros((N,), dtype=torch.int32, device=dev)
    tI = torch.zeros((N,), dtype=torch.int32, device=dev)
    gather = torch.zeros((world, N, n), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    # Collection of the moves on every rank (north_star: "all-gather only to collect ΔŨ").  Preferred: FUSED -- the
    # gather buffers live in symmetric memory (peer-mapped over NVLink) and the step kernel's epilogue stores Z̃ into
    # every peer's buffer; the only extra work per step is a device-side barrier.  Fallback: ncclAllGather.
    symm, gather_mode = None, "none"
    if world > 1:
        gather_mode = "ncclAllGather of Ztilde after every step"
        if os.environ.get("BMPC_FUSED_GATHER", "1") != "0":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                gsym = symm_mem.empty((world, N, n), dtype=torch.float64, device=dev)
                gsym.zero_()
                symm = symm_mem.rendezvous(gsym, dist.group.WORLD)
                b.set_gather([int(symm.buffer_ptrs[p]) for p in range(world)], rank)
                gather_mode = "fused: step-kernel epilogue stores Ztilde into every peer's symmetric-memory buffer (NVLink) + device barrier"
            except Exception as e:  # noqa: BLE001 -- any failure of the symmetric-memory path falls back to NCCL
                sys.stderr.write(f"[bench] symmetric memory unavailable ({e!r}); using ncclAllGather\n")
                symm = None

    def launch(k):
        b.step_device(dict(xhat0=tX[k].data_ptr(), lastu0=tLU[k].data_ptr(), ry=tRY[k].data_ptr(),
                           Ztilde=tZ[k].data_ptr(), u=tU.data_ptr(), J=tJ.data_ptr(), status=tS.data_ptr(),
                           iters=tI.data_ptr()))
        if world > 1:
            if symm is not None:
                symm.barrier()  # every peer's stores of this period have landed
            else:
                dist.all_gather_into_tensor(gather.view(-1), tZ[k].reshape(-1))

    tLU0, tZ0 = tLU.clone(), tZ.clone()
    BUSY_PASSES = 60
    for k in range(W):
        flush.zero_()
        launch(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    it_sum, st_bad = 0, 0
    l_before = b.launch_count()
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        t_wall0 = time.perf_counter()
        for j in range(K):
            flush.zero_()
            ev[j][0].record(stream)
            launch(W + j)
            ev[j][1].record(stream)
        torch.cuda.synchronize()
        t_wall1 = time.perf_counter()
        # extra passes without flushes keep the GPU busy long enough for the 10 Hz clock sampler
        # (never used for `value`); the in/out slices are restored from pristine copies before each pass
        for _ in range(BUSY_PASSES):
            tLU.copy_(tLU0)
            tZ.copy_(tZ0)
            for k in range(W, W + K):
                launch(k)
        torch.cuda.synchronize()
    launches_timed = (b.launch_count() - l_before) - BUSY_PASSES * K  # exclude the clock-sampling passes
    gather_check = None
    if world > 1 and symm is not None:
        # the fused gather against ncclAllGather on the last period's Z̃
        dist.all_gather_into_tensor(gather.view(-1), tZ[W + K - 1].reshape(-1))
        torch.cuda.synchronize()
        gather_check = bool(torch.equal(gather, gsym))
    dev_ms = sum(a.elapsed_time(c) for a, c in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    ms_per_step = dev_ms_max / K
    value = world * N * K / (dev_ms_max * 1e-3)
    iters_rec = rec["iters"][W:W + K]
    mean_iters = float(iters_rec.mean())
    frac_exit = float((iters_rec == 0).mean())
    n_nonopt = int((rec["status"][W:W + K] != 0).sum())

    # ---- e2e leg: the C-ABI call with pinned host buffers, H2D + kernel + D2H inside the timed region ----
    pin = lambda a: torch.from_numpy(a.copy()).pin_memory().numpy()
    hX, hLU, hRY, hZ, hY = pin(rec["xhat0"]), pin(rec["lastu0"]), pin(rec["ry"]), pin(rec["Zin"]), pin(rec["y0m"])
    # the observer (SteadyKalmanFilter of the recording run) moves into the step kernel: per period the host sends the
    # plant measurement ym and the setpoint, and reads back u -- exactly the arguments / result of the reference's
    # preparestate! + moveinput! + updatestate! sequence (src/plot_sim.jl:291-311)
    est = mpc.estim
    # BMPC_E2E_FUSED=1 moves the observer into the step kernel (ym instead of x̂0 goes up: 64 KiB less per period) --
    # measured 0.258 ms/step against 0.235 with the host-side observer (the two dependent matrix-vector products
    # lengthen every instance's start), so the default e2e leg keeps x̂0 as the input, as moveinput! has it.
    fused_obs = os.environ.get("BMPC_E2E_FUSED", "0") != "0"
    if fused_obs:
        b.set_estimator(est.Ahat, est.Buhat, est.Cmhat, est.Khat, None, None, est.fophat - est.xophat)
        b.set_state(np.zeros((N, b.nxhat)))
    hU = torch.zeros((N, nu), dtype=torch.float64).pin_memory().numpy()
    hJ = torch.zeros((N,), dtype=torch.float64).pin_memory().numpy()
    hS = torch.zeros((N,), dtype=torch.int32).pin_memory().numpy()
    hI = torch.zeros((N,), dtype=torch.int32).pin_memory().numpy()
    import ctypes as C
    from mpc_b200 import _lib

    # zero-copy: the pinned host buffers are device-accessible, the step kernel reads x̂0, ry from and writes u, status to
    # them directly over PCIe (io.host_mapped = 1; BMPC_E2E_ZEROCOPY=0 selects the library's explicit copies instead)
    zero_copy = int(os.environ.get("BMPC_E2E_ZEROCOPY", "1") != "0")

    def host_step(k, resident=1):
        # resident = 1: u0(k-1) and the previous Z̃ are state of the handle, as mpc.lastu0 / mpc.Z̃ are fields of the
        # reference LinMPC; x̂0 is state of the fused observer: per call ym, ry go up, u and the status come back.
        src = dict(y0m=hY[k].ctypes.data) if fused_obs else dict(xhat0=hX[k].ctypes.data)
        if resident:
            io = _lib.StepIO(ry=hRY[k].ctypes.data, u=hU.ctypes.data, status=hS.ctypes.data, device_ptrs=0, sync=1,
                             resident=1, host_mapped=zero_copy, **src)
        else:
            io = _lib.StepIO(lastu0=hLU[k].ctypes.data, ry=hRY[k].ctypes.data, Ztilde=hZ[k].ctypes.data, u=hU.ctypes.data,
                             J=hJ.ctypes.data, status=hS.ctypes.data, iters=hI.ctypes.data, device_ptrs=0, sync=1, **src)
        _lib.check(_lib.lib().bmpc_step(b._h, C.byref(io)))

    host_step(0, resident=0)  # loads the recorded u0(-1) / Z̃ of period 0 into the handle
    for k in range(1, W):
        host_step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(W, W + K):
        host_step(k)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_value = world * N * K / e2e_s
    h2d = N * 8 * ((ny if fused_obs else b.nxhat) + ny)
    d2h = N * (8 * nu + 4)
    # the replay reproduces the recorded inputs (it is not a closed loop: the recorded measurements do not react to
    # last-digit differences of u, so a few ill-conditioned instances drift when the observer runs on the device)
    du = np.abs(hU - rec["u"][W + K - 1]).max(axis=1)
    e2e_check = {"median_abs_du": float(np.median(du)), "frac_within_1e-6": float((du < 1e-6).mean())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    for f in ("MEASURED_PEAKS.json", os.path.join("profiles", "fp64_peak_r01.json")):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            peaks.update(json.load(open(p)))
    fp64_peak = peaks.get("fp64_tflops_used_as_peak", 35.48)
    m_ref = 2 * nu * Hp + ny * Hp + 1
    fl_step, fl_chol, f_asm, f_it = flops_model(nu, ny, b.nxhat, Hp, Hc, 1, m_ref, mean_iters)
    achieved = fl_step * N / (ms_per_step * 1e-3) / 1e12
    traffic = None
    ncu_sum = os.path.join(ROOT, "profiles", "ncu_r01_warp_summary.json")
    if os.path.exists(ncu_sum):
        traffic = json.load(open(ncu_sum)).get("dram_bytes_per_launch")
    # per-step I/O + per-instance constants (E, K, V, B, packed H, row bounds) in the reference's layout
    alg_bytes = N * 8 * (b.nxhat + 2 * nu + ny + 2 * n + 2) + N * 8 * (b.nY * (n + b.nxhat + nu + 1) + n * (n + 1) // 2 + m_ref)
    line = {
        "metric": "MPC steps/sec (batch LinMPC Hp=20 Hc=5 nu=2 ny=2)", "value": value, "unit": "instance-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "gpu_launches": int(launches_timed),
        "gather": {"mode": gather_mode, "equals_ncclAllGather": gather_check},
        "solver": {"mean_ipm_iters_per_instance": mean_iters, "unconstrained_exit_fraction": frac_exit,
                   "non_optimal_statuses": n_nonopt, "tol": 1e-11, "launch": b.launch_info()},
        "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA m8n8k4 + DFMA; B200's FP64 tensor rate equals its FP64 vector rate)", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                     "traffic": traffic, "peak_source": "profiles/fp64_peak_r01.json (cuBLAS DGEMM 8192^3 measured on this pool; "
                     "MEASURED_PEAKS.json has no fp64 entry)", "algorithmic_flops_per_instance_step": fl_step,
                     "cholesky_flops_per_instance_step": fl_chol, "achieved_cholesky_tflops": fl_chol * N / (ms_per_step * 1e-3) / 1e12,
                     "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                             "peak_gbs": peaks.get("hbm_gbs")}},
        "e2e": {"value": e2e_value, "unit": "instance-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / K, "copies_per_step": ("H2D ym, ry; D2H u, status (x̂0, u0(k-1), Z̃ are handle state: "
                "fused SteadyKalmanFilter + io.resident = 1)" if fused_obs else "H2D xhat0, ry; D2H u, status (u0(k-1), Z̃ "
                "are handle state, io.resident = 1)") + ("; zero-copy: the kernel reads / writes the pinned host buffers "
                "over PCIe (io.host_mapped = 1)" if zero_copy else "; cudaMemcpyAsync staging copies"), "u_vs_recorded_last_period": e2e_check},
        "clocks": clk.summary(),
        "wall_s_timed_region": t_wall1 - t_wall0,
    }
    if world == 1:
        threads = os.cpu_count() or 1
        n_inst = min(args.cpu_instances, N)
        res = cpu_leg(model, rec, (W, W + K), n_inst, threads, Hp, Hc)
        res1 = cpu_leg(model, rec, (W, W + min(K, 20)), min(n_inst, 128), 1, Hp, Hc)
        line["cpu_baseline"] = {"value": res["value"], "unit": "instance-steps/s", "cores": res["threads"], "kind": "port",
                                "sample": f"first {n_inst} of {N} instances x {K} recorded periods ({res['seconds']:.2f} s); "
                                          f"restated reference path = initpred!/linconstraint! + OSQP-style ADMM (OSQP binary "
                                          f"unavailable), mean {res['admm_iters_per_solve']:.1f} ADMM iterations/solve",
                                "one_core_value": res1["value"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
