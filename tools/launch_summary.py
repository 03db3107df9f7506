#!/usr/bin/env python
"""Per-kernel totals and shares from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python bench.py ...`).

    python tools/launch_summary.py gpurun_out/launches_X.csv [--skip-until-step]   > profiles/rN_launches_X.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)           # drop the argument list
    name = name.replace("sga::", "").replace("void ", "")
    return name[:70]


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((short(r["Kernel Name"]), us, r["Grid Size"], r["Block Size"]))
    tot = OrderedDict()
    for name, us, grid, block in rows:
        t = tot.setdefault(name, [0, 0.0, grid, block])
        t[0] += 1
        t[1] += us
    total = sum(t[1] for t in tot.values()) or 1.0
    print("| kernel | launches | total us | share | grid | block |")
    print("|---|---|---|---|---|---|")
    for name, (n, us, grid, block) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f%% | %s | %s |" % (name, n, us, 100.0 * us / total, grid, block))
    print("\n%d launches, %.1f us in total" % (len(rows), total))


if __name__ == "__main__":
    main()
