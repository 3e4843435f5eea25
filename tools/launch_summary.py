#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total time, share.

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/launches_rNN.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        tot[name][0] += 1
        tot[name][1] += ns
        rows.append(name)
    total = sum(v[1] for v in tot.values())
    print(f"# {path}: {len(rows)} launches, {total / 1e6:.3f} ms summed device time (cold-cache, serialised under ncu)")
    print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'us/launch':>10s}")
    for name, (cnt, ns) in sorted(tot.items(), key=lambda t: -t[1][1]):
        print(f"{name[:60]:60s} {cnt:8d} {ns / 1e6:10.3f} {ns / total:7.3f} {ns / 1e3 / cnt:10.1f}")


if __name__ == "__main__":
    main()
