#!/bin/bash
# definitive 1-GPU numbers of round 2 (final build): suite, both arms on smoke, ncu of smoke
O=gpurun_out/r2final
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -4 > $O/pytest_final.txt; cat $O/pytest_final.txt | tail -2
python bench.py --steps 20 --warmup 5 > $O/bench_fnx_n1.json 2> $O/bench_fnx_n1.err; echo "bench rc=$?"
python bench.py --frames-in-flight 1 --lanes 1 --no-cpu-baseline --no-dropin > $O/bench_fnx_n1_oneframe.json 2>/dev/null
for wl in scalar c2 ball; do
  python bench.py --workload $wl --steps 20 --warmup 5 > $O/bench_${wl}_n1.json 2> $O/bench_${wl}_n1.err
done
export FNX_WORKLOAD=smoke
bash tools/gpu_ncu.sh r2final/smoke blend_bwd_kernel blend_fwd_kernel > /dev/null 2>&1
timeout 300 python tools/profile_step.py > $O/smoke/profile_step.txt 2>&1
