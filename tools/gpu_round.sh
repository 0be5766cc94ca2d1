#!/bin/bash
# One GPU-box visit: GPU parity tests, both bench arms, ncu launch list + full captures of the top kernels.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-rX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt
python -c 'import torch; print(torch.cuda.get_device_name(0))'
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py > $OUT/bench_fnx_n1.json 2> $OUT/bench_fnx_n1.err; echo "bench rc=$?"; cat $OUT/bench_fnx_n1.json
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference > $OUT/bench_reference_n1.json 2> $OUT/bench_reference_n1.err; echo "ref rc=$?"; cat $OUT/bench_reference_n1.json
fi
timeout 300 python tools/profile_step.py > $OUT/profile_step.txt 2>&1; cat $OUT/profile_step.txt
if [ -z "$SKIP_NCU" ]; then
  FNX_ITERS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py > $OUT/launches.log 2>&1
  for k in blend_bwd_kernel blend_fwd_kernel ${NCU_EXTRA}; do
    FNX_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $OUT/ncu_$k python tools/profile_step.py > $OUT/ncu_$k.log 2>&1
  done
fi
ls -la $OUT
