#!/bin/bash
# ncu launch list of one iteration + full captures of the named kernels.  usage: bash tools/gpu_ncu.sh <tag> kernel...
TAG=${1:-rX}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
FNX_ITERS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py > $OUT/launches.log 2>&1
for k in "$@"; do
  FNX_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $OUT/ncu_$k python tools/profile_step.py > $OUT/ncu_$k.log 2>&1
done
ls -la $OUT
