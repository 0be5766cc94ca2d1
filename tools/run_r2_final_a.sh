#!/bin/bash
# round-2 final 1-GPU visit: full GPU suite, both bench arms on every workload, ncu launch lists + captures
O=gpurun_out/r2final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=8 > $O/pytest_full.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_full.txt
tail -3 $O/pytest_full.txt
grep -E "^c[2-5]|grad |merged streams|median-depth" $O/pytest_full.txt > $O/parity_baseline_sizes.txt
python bench.py --steps 20 --warmup 5 > $O/bench_fnx_n1.json 2> $O/bench_fnx_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err; echo "ref rc=$?"
for wl in scalar c2 ball; do
  python bench.py --workload $wl --steps 20 --warmup 5 > $O/bench_${wl}_n1.json 2> $O/bench_${wl}_n1.err
  python bench.py --impl reference --workload $wl --steps 10 --warmup 2 --ref-max-seconds 60 > $O/bench_reference_${wl}.json 2> $O/bench_reference_${wl}.err
done
python bench.py --frames-in-flight 1 --lanes 1 --no-cpu-baseline --no-dropin > $O/bench_fnx_n1_oneframe.json 2>/dev/null
for wl in smoke scalar c2 ball; do
  export FNX_WORKLOAD=$wl
  bash tools/gpu_ncu.sh r2final/$wl blend_bwd_kernel blend_fwd_kernel > /dev/null 2>&1
  timeout 300 python tools/profile_step.py > $O/$wl/profile_step.txt 2>&1
done
ls $O
