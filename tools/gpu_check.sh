#!/bin/bash
# quick GPU check used during kernel work: parity tests, then an un-profiled bench line.  usage: tools/gpu_check.sh TAG [pytest -k expr]
TAG=${1:-chk}
if [ -n "$2" ]; then
python -m pytest tests -m gpu -x -q -k "$2" > gpurun_out/${TAG}_pytest.log 2>&1
else
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
fi
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench.log 2>&1
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.log").read().strip().split("\n")[-1])
    k = d["roofline"].get("kernels", {})
    print("ms/step %.4f  value %.1fM  frac %.3f  e2e %.1fM (%.4f ms)  " % (d["ms_per_step"], d["value"] / 1e6, d["roofline"]["frac"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]) + "  ".join("%s %.4f ms frac %.3f" % (n, v["ms"], v["frac"]) for n, v in k.items()))
except Exception as ex:
    print("bench failed", ex); print(open("gpurun_out/${TAG}_bench.log").read()[-2000:])
PY
