"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for name, us in rows:
        short = re.sub(r"<.*", "", name)[:70]
        agg[short][0] += 1
        agg[short][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {total:.1f} us total (serialised, cold-cache: compare shares)")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{us:10.1f} us {100 * us / total:5.1f}%  x{n:<4d} {name}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
