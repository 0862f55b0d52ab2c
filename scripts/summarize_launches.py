"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=25):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name)
        rows.append((name, ns))
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {total / 1e6:.2f} ms summed kernel time")
    print(f"{'kernel':70s} {'count':>6s} {'ms':>9s} {'share':>7s} {'avg us':>9s}")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{n[:70]:70s} {c:6d} {ns / 1e6:9.3f} {100 * ns / total:6.1f}% {ns / c / 1e3:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
