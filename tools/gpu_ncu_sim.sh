#!/bin/bash
# full capture of the simulation kernel only (steady state).  usage: tools/gpu_ncu_sim.sh TAG
TAG=${1:-prof}
CMD="python bench.py --steps 40 --warmup 5 --burn-in 300 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 322 -c 1 -f -o gpurun_out/${TAG}_sim $CMD > gpurun_out/${TAG}_full_sim.log 2>&1
tail -2 gpurun_out/${TAG}_full_sim.log
