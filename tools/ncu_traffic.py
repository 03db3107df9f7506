#!/usr/bin/env python
"""DRAM traffic of one step's tensor kernels from an `ncu --set full` report -> profiles/<tag>_traffic.json (read by
bench.py for `roofline.traffic`).

    python tools/ncu_traffic.py gpurun_out/prof_X.ncu-rep profiles/r1_X_traffic.json "<how the report was captured>"

The report must hold the 14 launches of one step in launch order (bench.py under
`ncu -k regex:"k_conv|k_l3" -s 42 -c 14`): hs2/hs1/hs0 dgrad of the previous step, hs0-2 fwd, gs0-2 fwd, l3 fwd,
l3 dgrad, gs2/gs1/gs0 dgrad.
"""
import csv
import io
import json
import subprocess
import sys

LABELS = ["hs2_dgrad*", "hs1_dgrad*", "hs0_dgrad*", "hs0_fwd", "hs1_fwd", "hs2_fwd", "gs0_fwd", "gs1_fwd", "gs2_fwd",
          "l3_fwd", "l3_dgrad_igdn", "gs2_dgrad", "gs1_dgrad", "gs0_dgrad"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out_path, how = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def nbytes(r, k):
        return float(r[col[k]]) * SCALE[units[col[k]]]

    res = []
    for lab, r in zip(LABELS, data):
        res.append(dict(label=lab, kernel=r[col["Kernel Name"]][:40], time=float(r[col["gpu__time_duration.sum"]]),
                        time_unit=units[col["gpu__time_duration.sum"]],
                        dram_read_bytes=nbytes(r, "dram__bytes_read.sum"), dram_write_bytes=nbytes(r, "dram__bytes_write.sum"),
                        tensor_pipe_pct=float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])))
    conv = [x for x in res if not x["label"].startswith("l3_")]
    tot = sum(x["dram_read_bytes"] + x["dram_write_bytes"] for x in conv)
    j = dict(source=how,
             note="(de)conv kernels of one step: 12 launches (k_conv_tc x9, k_conv_igdn_fwd x3); rows marked * are the hs "
                  "dgrads of the neighbouring step",
             conv_launches=len(conv), conv_dram_bytes_per_step=tot, conv_dram_bytes_per_launch=tot / max(1, len(conv)),
             kernels=res)
    with open(out_path, "w") as fh:
        json.dump(j, fh, indent=1)
    print("%.3f GB over %d launches" % (tot / 1e9, len(conv)))


if __name__ == "__main__":
    main()
