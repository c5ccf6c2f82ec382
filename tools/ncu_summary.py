"""Summarise an ncu --csv launch list (gpu__time_duration.sum): per kernel name (+ grid) total time, launches, share.
Usage: python tools/ncu_summary.py launches.csv [top]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1e-3)
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        rows.append((name, r.get("Grid Size", ""), v))
    agg = defaultdict(lambda: [0.0, 0])
    for name, grid, us in rows:
        agg[name][0] += us
        agg[name][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"{len(rows)} launches, {total / 1e3:.3f} ms under ncu (cold-cache, serialised)")
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{us / 1e3:9.3f} ms {100 * us / total:5.1f}%  x{n:<4d} avg {us / n:8.1f} us  {name[:110]}")


if __name__ == "__main__":
    main()
