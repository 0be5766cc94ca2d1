#!/bin/bash
# round 2, second session: definitive 1-GPU numbers of the final build -- suite, bench on every workload, ncu launch lists + captures
O=gpurun_out/s2final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=8 > $O/pytest_full.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_full.txt
tail -3 $O/pytest_full.txt
grep -E "^c[2-5]|grad |merged streams|median-depth" $O/pytest_full.txt > $O/parity_baseline_sizes.txt
python bench.py --steps 20 --warmup 5 > $O/bench_fnx_n1.json 2> $O/bench_fnx_n1.err; echo "bench rc=$?"
python bench.py --frames-in-flight 1 --lanes 1 --no-cpu-baseline --no-dropin > $O/bench_fnx_n1_oneframe.json 2>/dev/null
for wl in scalar c2 ball; do
  python bench.py --workload $wl --steps 20 --warmup 5 > $O/bench_${wl}_n1.json 2> $O/bench_${wl}_n1.err
done
for wl in smoke scalar c2 ball; do
  export FNX_WORKLOAD=$wl
  EXTRA=""
  if [ $wl = smoke ]; then EXTRA="density_bwd_kernel merge_bucket_kernel advect_fwd_kernel"; fi
  bash tools/gpu_ncu.sh s2final/$wl blend_bwd_kernel blend_fwd_kernel $EXTRA > /dev/null 2>&1
  timeout 300 python tools/profile_step.py > $O/$wl/profile_step.txt 2>&1
done
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
ls $O
