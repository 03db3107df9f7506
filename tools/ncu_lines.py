#!/usr/bin/env python
"""Top stall lines (SASS) of one kernel of an .ncu-rep captured with --import-source on, with the dominant stall reasons.

    python tools/ncu_lines.py rep.ncu-rep <kernel name substring> [top N]
"""
import csv
import io
import subprocess
import sys


def blocks(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    res, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            res.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = row
            else:
                cur["rows"].append(row)
    return res


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    b = [x for x in blocks(rep) if pat in x["name"]][0]
    h = {k: i for i, k in enumerate(b["hdr"])}
    si = h["# Samples"]
    tot = sum(int(r[si] or 0) for r in b["rows"]) or 1
    st = [k for k in b["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[h[k]] or 0) for r in b["rows"]) for k in st}
    print(b["name"], "samples", tot)
    print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
    ranked = sorted(enumerate(b["rows"]), key=lambda t: -int(t[1][si] or 0))[:top]
    for n, r in sorted(ranked):
        why = sorted(((int(r[h[k]] or 0), k[6:]) for k in st), reverse=True)[:2]
        print("%5d %5.1f%% %9s  %-64s %s" % (n, 100.0 * int(r[si] or 0) / tot, r[h["Instructions Executed"]],
                                            r[h["Source"]].strip()[:64], " ".join("%s:%d" % (k, v) for v, k in why if v)))


if __name__ == "__main__":
    main()
