#!/bin/bash
# A/B of scheduling switches on the GPU box: prints value / e2e it/s of bench.py for each environment setting.
# usage: tools/ab_bench.sh <tag> "<bench args>" "ENV1=.. ENV2=.." "ENV3=.." ...   (use "-" for the default environment)
TAG=$1; shift; ARGS=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  [ "$envs" = "-" ] && envs=""
  env $envs timeout 300 python bench.py $ARGS --no-cpu > gpurun_out/ab_${TAG}_$i.log 2>&1
  python - "$envs" gpurun_out/ab_${TAG}_$i.log <<'PY'
import json, sys
env, path = sys.argv[1], sys.argv[2]
line = [l for l in open(path) if l.startswith("{")]
if not line:
    print("%-40s FAILED (see %s)" % (env or "(default)", path)); sys.exit(0)
d = json.loads(line[-1])
k = d["roofline"]["kernels"]
print("%-40s value %8.2f it/s (%.4f ms)  e2e %8.2f  gs2_fwd %.4f ms  elementwise %.4f  instrumented-sum %.3f" % (
    env or "(default)", d["value"], d["ms_per_step"], d["e2e"]["value"], k.get("gs2_fwd", {}).get("ms_per_step", 0),
    k.get("elementwise", {}).get("ms_per_step", 0), sum(v["ms_per_step"] for v in k.values())))
PY
done
