#!/bin/bash
O=gpurun_out/s2g; mkdir -p $O
for wl in smoke scalar; do
FNX_WORKLOAD=$wl FNX_ITERS=2 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file $O/launches_$wl.csv python tools/profile_step.py > $O/launches_$wl.log 2>&1
done
SKIP_TESTS=1 bash tools/exp_ab.sh s2g smoke
