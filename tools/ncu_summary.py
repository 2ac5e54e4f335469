#!/usr/bin/env python
"""Text summary of an `ncu --set full` report for profiles/: python tools/ncu_summary.py <rep> > profiles/<name>.txt
Key raw metrics of the first kernel in the report + per-function instruction/stall shares (needs -lineinfo)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u, v = r[0], r[1], r[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio"]
print(f"# {rep}")
for a, b, c in zip(h, u, v):
    if a in want or ("issue_stalled" in a and "per_issue_active" in a):
        print(f"{a:90s} {b:12s} {c}")
print()
sys.stdout.flush()
subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "25"], stdout=subprocess.DEVNULL)
print("# per-function share (source lines bucketed by the kg_* function that contains them)")
sys.stdout.flush()
subprocess.run([sys.executable, "tools/ncu_buckets.py"])
