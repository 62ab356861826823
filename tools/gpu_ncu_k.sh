#!/bin/bash
# full capture of one kernel (regex) at steady state, warm caches.  usage: tools/gpu_ncu_k.sh TAG KERNEL_REGEX [skip]
TAG=${1:-prof}; K=${2:-env_kernel}; SKIP=${3:-600}
CMD="python bench.py --steps 40 --warmup 5 --burn-in 300 --no-cpu-baseline"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/${TAG}_$K $CMD > gpurun_out/${TAG}_full_$K.log 2>&1
tail -2 gpurun_out/${TAG}_full_$K.log
