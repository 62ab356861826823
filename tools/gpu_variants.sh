#!/bin/bash
# A/B of prebuilt library variants (build/var/lib_<name>.so, e.g. compiled with -DAT_BULLET_SPLIT=1): un-profiled bench
# line + oracle parity subset per variant.  usage: tools/gpu_variants.sh TAG name1 name2 ...   (base = the in-tree build)
TAG=$1; shift
cp alphatank_b200/libalphatank_b200.so /tmp/lib_base.so
one() {
    python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_$1.log 2>&1
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_$1.log").read().strip().split("\n")[-1])
    k = d["roofline"].get("kernels", {})
    print("%-8s ms/step %.4f  value %.1fM  e2e %.1fM  " % ("$1", d["ms_per_step"], d["value"] / 1e6, d["e2e"]["value"] / 1e6) + " ".join("%s %.4f" % (n[:4], v["ms"]) for n, v in k.items()))
except Exception as ex:
    print("$1 bench failed", ex); print(open("gpurun_out/${TAG}_bench_$1.log").read()[-1500:])
PY
}
one base
for v in "$@"; do
    cp build/var/lib_$v.so alphatank_b200/libalphatank_b200.so
    one $v
    [ -n "$SKIP_TESTS" ] || timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "oracle_on_batches or full_size or specialisations" > gpurun_out/${TAG}_pytest_$v.log 2>&1
    tail -2 gpurun_out/${TAG}_pytest_$v.log
done
cp /tmp/lib_base.so alphatank_b200/libalphatank_b200.so
one base2
