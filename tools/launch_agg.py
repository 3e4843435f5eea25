#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into one line per kernel.

    python tools/launch_agg.py launches.csv [--last-half]

--last-half keeps the second half of the launches (a warm-up call followed by an identical timed call)."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("Grid Size")
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[ui], 1.0)
        name = r[ki].split("(")[0].split("::")[-1]
        seq.append((name, v))
    if "--last-half" in sys.argv:
        seq = seq[len(seq) // 2:]
    agg = collections.OrderedDict()
    for name, v in seq:
        agg.setdefault(name, []).append(v)
    tot = sum(v for _, v in seq)
    print(f"{len(seq)} launches, {tot / 1e3:.2f} ms in kernels (ncu: serialised, --cache-control none)")
    print(f"{'kernel':34s} {'launches':>8s} {'total ms':>9s} {'share':>6s} {'mean us':>8s} {'min us':>8s} {'max us':>8s}")
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{name:34s} {len(v):8d} {sum(v) / 1e3:9.2f} {sum(v) / tot:6.1%} {sum(v) / len(v):8.2f} {min(v):8.2f} {max(v):8.2f}")


if __name__ == "__main__":
    main()
