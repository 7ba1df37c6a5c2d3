"""Per-kernel shares of a window of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file ...`):
    python scripts/launch_shares.py launches.csv[.gz] [first_id last_id]
Serialised, cold-cache durations: compare SHARES, not absolutes (B200_PROFILING.md)."""
import csv
import gzip
import io
import sys
from collections import defaultdict

path = sys.argv[1]
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
lines = [ln for ln in raw.splitlines() if ln.startswith('"')]
rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
acc = defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum" or not (lo <= int(r["ID"]) <= hi):
        continue
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
    name = r["Kernel Name"].split("(")[0].replace("void ", "")
    key = (name, r["Grid Size"].replace(" ", ""))
    acc[key][0] += 1
    acc[key][1] += v
    tot += v
print(f"# {path}: launches {lo}..{hi if hi < 10**9 else 'end'}, {sum(a[0] for a in acc.values())} launches, {tot / 1e3:.2f} ms of kernel time")
for (name, grid), (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:28s} grid={grid:>14s} launches={n:6d} total_us={us:11.1f} avg_us={us / n:9.2f} share={us / tot:.3f}")
