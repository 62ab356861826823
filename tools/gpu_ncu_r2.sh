#!/bin/bash
# round 2: launch list of the bench command + full captures of the step's two kernels and of the policy kernel, each after
# the plain command has exited 0.  usage: tools/gpu_ncu_r2.sh TAG
TAG=${1:-r06}
CMD="python bench.py --steps 40 --warmup 5 --burn-in 300 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain run failed; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rows_kernel -s 320 -c 1 -f -o gpurun_out/${TAG}_rows $CMD > gpurun_out/${TAG}_full_rows.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 322 -c 1 -f -o gpurun_out/${TAG}_sim $CMD > gpurun_out/${TAG}_full_sim.log 2>&1
PCMD="python tools/policy_probe.py"
$PCMD > gpurun_out/${TAG}_policy_plain.log 2>&1 || { echo policy probe failed; tail -20 gpurun_out/${TAG}_policy_plain.log; exit 1; }
# policy_probe launches 35 x {fp32 rows, all chunks | fp32 rows, folded | fp16 rows, all chunks | fp16 rows, folded}: launch 120 = fp16 folded, 50 = fp32 folded
ncu --set full --clock-control none --import-source on -k regex:policy_forward_kernel -s 120 -c 1 -f -o gpurun_out/${TAG}_policy16 $PCMD > gpurun_out/${TAG}_full_policy16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy_forward_kernel -s 50 -c 1 -f -o gpurun_out/${TAG}_policy32 $PCMD > gpurun_out/${TAG}_full_policy32.log 2>&1
RCMD="python tools/rollout_probe.py --half --no-graph --iters 2"
$RCMD > gpurun_out/${TAG}_rollout_plain.log 2>&1 || { echo rollout probe failed; tail -20 gpurun_out/${TAG}_rollout_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 150 --csv --log-file gpurun_out/${TAG}_rollout_launches.csv $RCMD > gpurun_out/${TAG}_rollout_ncu.log 2>&1
ls -la gpurun_out/${TAG}_*
