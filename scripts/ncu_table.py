"""Compact per-launch table from the raw-page CSV of an ncu --set full report (gzipped or not):
    python scripts/ncu_table.py gpurun_out/r2_step_l3_raw.csv.gz
kernel, grid, duration, DRAM bytes (read + write), DRAM / L2 / SM throughput %, FP64 pipe %, issue %, warps active %,
registers, instructions per thread, top stall reasons."""
import csv
import gzip
import io
import sys

path = sys.argv[1]
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def num(d, k, default=float("nan")):
    try:
        return float(d[col[k]].replace(",", ""))
    except Exception:
        return default


def unit(k):
    return units[col[k]] if k in col else ""


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


print(f"{'kernel':34s} {'grid':>7s} {'us':>9s} {'dramMB':>9s} {'GB/s':>7s} {'dram%':>6s} {'l2%':>5s} {'sm%':>5s} {'fp64%':>6s} "
      f"{'issue%':>6s} {'warps%':>6s} {'regs':>4s} {'inst/thr':>8s}  stalls")
for d in data:
    name = d[col["Kernel Name"]].split("(")[0].replace("void ", "")
    grid = d[col["launch__grid_size"]] if "launch__grid_size" in col else "?"
    us = to_us(num(d, "gpu__time_duration.sum"), unit("gpu__time_duration.sum"))
    rd = to_bytes(num(d, "dram__bytes_read.sum"), unit("dram__bytes_read.sum"))
    wr = to_bytes(num(d, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"))
    threads = num(d, "launch__thread_count", 0) or (num(d, "launch__grid_size", 0) * num(d, "launch__block_size", 0))
    inst = num(d, "smsp__inst_executed.sum", 0) * 32 / max(threads, 1)
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(d[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print(f"{name[:34]:34s} {grid:>7s} {us:9.1f} {(rd + wr) / 1e6:9.1f} {(rd + wr) / us / 1e3:7.0f} "
          f"{num(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{num(d, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"{num(d, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"{num(d, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{num(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{num(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{d[col['launch__registers_per_thread']] if 'launch__registers_per_thread' in col else '?':>4s} {inst:8.0f}  "
          + " ".join(f"{n}={v:.1f}" for v, n in stalls[:4]))
