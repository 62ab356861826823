#!/bin/bash
# launch list (per-kernel durations) of the bench at steady state.  usage: tools/gpu_launches.sh TAG
TAG=${1:-prof}
CMD="python bench.py --steps 40 --warmup 5 --burn-in 300 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1800 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%4d x %-60s avg %8.1f us  %5.1f %%" % (len(v), k, sum(v) / len(v), 100 * sum(v) / tot))
PY
