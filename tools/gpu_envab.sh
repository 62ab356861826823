#!/bin/bash
# A/B of runtime switches (environment variables) on the in-tree build.  usage: tools/gpu_envab.sh TAG "VAR=1" "VAR2=1 VAR3=2" ...
TAG=$1; shift
one() {  # name, env assignments
    local name=$1; shift
    env "$@" python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_${name}.log 2>&1
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.log").read().strip().split("\n")[-1])
    k = d["roofline"].get("kernels", {})
    print("%-10s ms/step %.4f  value %.1fM  e2e %.1fM  " % ("${name}", d["ms_per_step"], d["value"] / 1e6, d["e2e"]["value"] / 1e6) + " ".join("%s %.4f" % (n[:4], v["ms"]) for n, v in k.items()))
except Exception as ex:
    print("${name} bench failed", ex); print(open("gpurun_out/${TAG}_bench_${name}.log").read()[-1500:])
PY
}
one base X=1
i=0
for v in "$@"; do i=$((i+1)); one "v$i" $v; done
one base2 X=1
