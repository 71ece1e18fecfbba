"""profiles/latest.json from an `ncu --page raw --csv` export of the step kernel's timed bench launch.
usage: python scripts/ncu_summary.py RAW.csv SOURCE_NOTE STEPS_PER_LAUNCH REPLICAS > profiles/latest.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
name, unit, val = rows[0], rows[1], rows[2]
col = {n: i for i, n in enumerate(name)}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}


def get(metric, scaled=False):
    v = float(val[col[metric]])
    return v * SCALE.get(unit[col[metric]], 1.0) if scaled else v


steps, replicas = int(sys.argv[3]), int(sys.argv[4])
rd, wr = get("dram__bytes_read.sum", True), get("dram__bytes_write.sum", True)
out = {
    "source": sys.argv[2],
    "steps_per_launch": steps,
    "replicas": replicas,
    "dram_bytes_read": rd,
    "dram_bytes_write": wr,
    "dram_bytes_per_launch": rd + wr,
    "dram_bytes_per_step": (rd + wr) / steps,
    "duration_ms_under_ncu": get("gpu__time_duration.sum", True),
    "issue_slot_utilization": get("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
    "ipc": get("sm__inst_executed.avg.per_cycle_active"),
    "warp_instructions_per_replica_step": get("smsp__inst_executed.sum") / (replicas * steps * 200.0),
    "achieved_occupancy_warps_of_64": get("sm__warps_active.avg.per_cycle_active"),
    "smem_wavefronts_pct_of_peak": get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    "l2_hit_rate_pct": get("lts__t_sector_hit_rate.pct"),
    "eligible_warps_per_cycle": get("smsp__warps_eligible.avg.per_cycle_active"),
    "threads_per_warp_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "stalls_per_issue": {k: get(f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio")
                         for k in ("wait", "short_scoreboard", "long_scoreboard", "not_selected", "barrier", "branch_resolving", "math_pipe_throttle")},
    "pipe_fma_pct": get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "pipe_alu_pct": get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "pipe_fp64_pct": get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
}
print(json.dumps(out, indent=1))
