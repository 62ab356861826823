#!/bin/bash
# A/B of the split (simulation + rows_kernel) and fused step paths.  usage: tools/gpu_split.sh TAG
TAG=${1:-split}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_rows or masked_reset or step_host" > gpurun_out/${TAG}_pytest_split.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_split.log; fi
one() {  # name, env assignments...
    local name=$1; shift
    env "$@" timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_${name}.log 2>&1
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.log").read().strip().split("\n")[-1])
    k = d["roofline"].get("kernels", {})
    print("%-8s ms/step %.4f  value %.1fM  frac %.3f  e2e %.1fM (%.4f ms)  " % ("${name}", d["ms_per_step"], d["value"] / 1e6, d["roofline"]["frac"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]) + " ".join("%s %.4f" % (n[:4], v["ms"]) for n, v in k.items()))
except Exception as ex:
    print("${name} bench failed", ex); print(open("gpurun_out/${TAG}_bench_${name}.log").read()[-1500:])
PY
}
one g2w8 X=1
one fused AT_FUSED_ROWS=1
one g4w11 AT_ROWS_LGG=2 AT_ROWS_WARPS=11
one g2w2 AT_ROWS_WARPS=2
one g2w16 AT_ROWS_WARPS=16
one nopdl AT_NO_PDL=1
