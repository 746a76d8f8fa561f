#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (+ template args for the GEMM)."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        m = re.match(r"(?:void )?(?:mm::)?(\w+)(<[^>]*>)?", name)
        key = (m.group(1) + (m.group(2) or "")) if m else name
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':58s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:58s} {n:8d} {ms:10.3f} {100 * ms / tot:6.1f}%")
    print(f"{'TOTAL':58s} {sum(v[0] for v in agg.values()):8d} {tot:10.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
