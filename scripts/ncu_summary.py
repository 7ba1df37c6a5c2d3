"""Print a compact summary of an .ncu-rep (key throughput metrics + warp stall reasons per launch)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
        "smsp__inst_executed.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__maximum_warps_per_active_cycle_pct"]
for d in data:
    print("-" * 100)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:75s} {d[i]:>18s} {units[i]}")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(d[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("stall reasons (warps stalled per issue):", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:7]))
