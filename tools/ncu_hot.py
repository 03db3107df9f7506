#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics, and for one kernel the SASS lines with the most stall samples.

    python tools/ncu_hot.py rep.ncu-rep                 # table of launches
    python tools/ncu_hot.py rep.ncu-rep <launch index>  # top stall lines of that launch (needs -lineinfo + --import-source)
"""
import csv
import io
import subprocess
import sys


TIME_MS = {"s": 1e3, "second": 1e3, "ms": 1.0, "msecond": 1.0, "us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6}
BYTES_GB = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}


def raw(rep, with_units=False):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if with_units:
        return rows[0], rows[1], rows[2:]
    return rows[0], rows[2:]


def table(rep):
    hdr, units, data = raw(rep, with_units=True)
    col = {h: i for i, h in enumerate(hdr)}
    def g(r, k):
        if k not in col:
            return "-"
        v, u = r[col[k]], units[col[k]]
        if k == "gpu__time_duration.sum":
            return float(v) * TIME_MS.get(u, 1.0)          # -> ms
        if k.startswith("dram__bytes"):
            return float(v) * BYTES_GB.get(u, 1.0)         # -> GB
        return v
    print("| # | kernel | ms | tensor pipe % | DRAM rd GB | DRAM wr GB | DRAM % | L2 hit % | regs | top stalls |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for n, r in enumerate(data):
        st = []
        for h, i in col.items():
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    st.append((float(r[i]), h.split("stalled_")[1]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        top = ", ".join("%s %.0f%%" % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:3])
        name = g(r, "Kernel Name").replace("(int)", "").replace("sga::", "")[:40]
        print("| %d | %s | %.4f | %.1f | %.3f | %.3f | %.1f | %.1f | %s | %s |" % (
            n, name, float(g(r, "gpu__time_duration.sum")),
            float(g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or 0),
            float(g(r, "dram__bytes_read.sum")), float(g(r, "dram__bytes_write.sum")),
            float(g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
            float(g(r, "lts__t_sector_hit_rate.pct")), g(r, "launch__registers_per_thread"), top))


def hot(rep, idx, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    # the output is a concatenation of per-kernel tables, each introduced by a "Kernel Name" row
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = row
            else:
                cur["rows"].append(row)
    b = blocks[idx * (len(blocks) // max(1, len(raw(rep)[1])))]
    h = {k: i for i, k in enumerate(b["hdr"])}
    si = h["# Samples"]
    tot = sum(int(r[si] or 0) for r in b["rows"]) or 1
    print(b["name"], "samples", tot)
    ranked = sorted(enumerate(b["rows"]), key=lambda t: -int(t[1][si] or 0))[:top]
    for n, r in sorted(ranked):
        print("%5d %5.1f%%  %s" % (n, 100.0 * int(r[si] or 0) / tot, r[h["Source"]].strip()[:110]))


if __name__ == "__main__":
    if len(sys.argv) == 2:
        table(sys.argv[1])
    else:
        hot(sys.argv[1], int(sys.argv[2]))
