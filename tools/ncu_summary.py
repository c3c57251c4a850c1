"""Summary of one kernel of an `ncu --set full` report as JSON (what profiles/ncu_*_summary.json hold):
    python tools/ncu_summary.py gpurun_out/warp_r02d.ncu-rep > profiles/ncu_r02_warp_summary.json
Reads the report with `ncu -i ... --page raw --csv` (ncu is in the image; no GPU needed)."""
import csv, io, json, subprocess, sys


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[2]
    d = dict(zip(hdr, vals))

    def f(k):
        try:
            return float(d[k])
        except (KeyError, ValueError):
            return None

    def scaled(k):  # byte metrics come with a unit (Kbyte / Mbyte ...)
        v, u = f(k), dict(zip(hdr, rows[1])).get(k, "")
        if v is None:
            return None
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    stall = {}
    for k in hdr:
        p, s = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
        if k.startswith(p) and k.endswith(s) and f(k) and f(k) >= 0.02 and "selected_per" not in k[len(p) - 1:len(p) + 9]:
            stall[k[len(p):-len(s)]] = round(f(k), 3)
    rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
    out = {
        "kernel": d.get("Kernel Name"),
        "report": rep,
        "duration_us": f("gpu__time_duration.sum"),
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": (rd or 0) + (wr or 0),
        "warp_instructions": f("smsp__inst_executed.sum"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct_of_peak": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_active_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "dmma_pipe_pct": f("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": f("launch__registers_per_thread"),
        "grid": f("launch__grid_size"), "block": f("launch__block_size"),
        "ctas_per_sm_limit_registers": f("launch__occupancy_limit_registers"),
        "ctas_per_sm_limit_smem": f("launch__occupancy_limit_shared_mem"),
        "dyn_smem_per_cta_bytes": scaled("launch__shared_mem_dynamic_per_block") if "launch__shared_mem_dynamic_per_block" in d else None,
        "eligible_warps_per_cycle": f("smsp__warps_eligible.avg.per_cycle_active"),
        "warp_latency_per_instruction_cycles": f("smsp__average_warp_latency_per_inst_issued.ratio"),
        "dram_throughput_pct_of_peak": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "stall_per_issue": dict(sorted(stall.items(), key=lambda kv: -kv[1])),
        "active_threads_per_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
