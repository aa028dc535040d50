"""Prints the metrics we track from an .ncu-rep (read on the CPU box): python tools/ncu_summary.py file.ncu-rep [row]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_tensor.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
        "lts__t_bytes.sum", "sm__sass_inst_executed_op_shared_ld.sum", "sm__sass_inst_executed_op_shared_st.sum",
        "sm__sass_inst_executed_op_global_ld.sum", "sm__sass_inst_executed_op_global_st.sum", "smsp__warps_eligible.avg.per_cycle_active"]
for r in rows[2:]:
    print("# ----")
    for h, u, v in zip(hdr, units, r):
        if h in KEEP or "issue_stalled" in h and "pct" in h and "not_issued" not in h:
            print("%s [%s] = %s" % (h, u, v))
