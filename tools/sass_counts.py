"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA in libsga_b200.so (B200_PROFILING.md):
UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM (tcgen05.ld), UTMALDG (TMA tensor load), UTMAPF (TMA L2 prefetch),
UTCBAR (tcgen05.commit), SYNCS (mbarrier).  Usage: python tools/sass_counts.py > profiles/r2_sass_counts.md"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "improving-inference-for-neural-image-compression_b200", "libsga_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = {}
        continue
    if cur is None:
        continue
    for mn in ("UTCHMMA", "LDTM", "UTMALDG", "UTMAPF", "UTCBAR", "SYNCS", "STTM"):
        if re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", line):
            kernels[cur][mn] = kernels[cur].get(mn, 0) + 1
            if mn == "UTCHMMA" and ".2CTA" in line:
                kernels[cur]["UTCHMMA.2CTA"] = kernels[cur].get("UTCHMMA.2CTA", 0) + 1
dem = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(kernels, dem)) if len(dem) == len(kernels) else {k: k for k in kernels}
print("# SASS evidence: tensor-core / TMEM / TMA instructions per kernel of libsga_b200.so (`cuobjdump -sass`, sm_100a)\n")
print("| kernel | UTCHMMA (of which .2CTA) | LDTM | UTMALDG | UTMAPF | UTCBAR | SYNCS |")
print("|---|---|---|---|---|---|---|")
tot = {}
for k, c in sorted(kernels.items(), key=lambda kv: names[kv[0]]):
    if not any(c.get(m) for m in ("UTCHMMA", "LDTM", "UTMALDG")):
        continue
    nm = re.sub(r"\([^()]*\)$", "", names[k]).replace("void ", "").replace("sga::", "").replace("(int)", "")
    print("| `%s` | %d (%d) | %d | %d | %d | %d | %d |" % (nm, c.get("UTCHMMA", 0), c.get("UTCHMMA.2CTA", 0), c.get("LDTM", 0),
                                                        c.get("UTMALDG", 0), c.get("UTMAPF", 0), c.get("UTCBAR", 0), c.get("SYNCS", 0)))
    for m, v in c.items():
        tot[m] = tot.get(m, 0) + v
print("\nTotals: " + ", ".join("%s %d" % kv for kv in sorted(tot.items())))
print("\nKernels without tensor-core instructions (elementwise, fp32 SIMT init / eval, codec, packing): %d"
      % sum(1 for c in kernels.values() if not any(c.get(m) for m in ("UTCHMMA", "LDTM", "UTMALDG"))))
