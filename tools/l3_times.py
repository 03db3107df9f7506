"""Prints the per-kernel times of bench.py JSON lines (development aid)."""
import json
import sys
for p in sys.argv[1:]:
    try:
        l = [x for x in open(p) if x.startswith("{")]
        d = json.loads(l[-1]); k = d["roofline"]["kernels"]
        print(p, "value %.2f" % d["value"], {n: (v["ms_per_step"], v.get("achieved")) for n, v in k.items()
                                             if n.startswith("l3") or n in ("gs0_fwd", "gs1_fwd", "gs2_fwd", "gs2_dgrad", "elementwise", "igdn_fwd", "igdn_bwd")})
    except Exception as e:
        print(p, "failed:", e)
