#!/bin/bash
# r2: full GPU test suite, then per-workload ncu launch lists + full captures of the blend kernels
mkdir -p gpurun_out/r2e
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -40 > gpurun_out/r2e/pytest.txt
tail -4 gpurun_out/r2e/pytest.txt
for wl in smoke scalar c2 ball; do
  export FNX_WORKLOAD=$wl
  bash tools/gpu_ncu.sh r2e/$wl blend_bwd_kernel blend_fwd_kernel > /dev/null 2>&1
  timeout 300 python tools/profile_step.py > gpurun_out/r2e/$wl/profile_step.txt 2>&1
  tail -3 gpurun_out/r2e/$wl/profile_step.txt
done
FNX_WORKLOAD=smoke FNX_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:density_bwd_kernel -s 8 -c 1 -f -o gpurun_out/r2e/smoke/ncu_density_bwd_kernel python tools/profile_step.py > /dev/null 2>&1
FNX_WORKLOAD=smoke FNX_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_bucket_kernel -s 4 -c 1 -f -o gpurun_out/r2e/smoke/ncu_merge_bucket_kernel python tools/profile_step.py > /dev/null 2>&1
du -sh gpurun_out/r2e
