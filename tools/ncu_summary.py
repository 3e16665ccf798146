"""Key metrics of an .ncu-rep (one kernel) as text. usage: ncu_summary.py rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "smsp__cycles_active.avg"]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for i, h in enumerate(hdr):
        if h in want or (h.startswith("smsp__average_warp") and "per_issue_active" in h and "not_issued" not in h) or \
           (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
            print("  %-75s %-12s %s" % (h, units[i], r[i]))
