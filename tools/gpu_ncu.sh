#!/bin/bash
# launch list + full captures of the step's kernels at steady state.  usage: tools/gpu_ncu.sh TAG
TAG=${1:-prof}
CMD="python bench.py --steps 40 --warmup 5 --burn-in 300 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain run failed; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rows_kernel -s 320 -c 1 -f -o gpurun_out/${TAG}_rows $CMD > gpurun_out/${TAG}_full_rows.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 322 -c 1 -f -o gpurun_out/${TAG}_sim $CMD > gpurun_out/${TAG}_full_sim.log 2>&1
tail -3 gpurun_out/${TAG}_full_sim.log
