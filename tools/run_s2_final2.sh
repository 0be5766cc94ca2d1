#!/bin/bash
# round 2, last build: suite + headline bench line + one-frame line + refreshed ncu launch list / captures of the kernels that changed last
O=gpurun_out/s2final2
mkdir -p $O/smoke
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -4 > $O/pytest_final.txt; tail -2 $O/pytest_final.txt
python bench.py --steps 20 --warmup 5 > $O/bench_fnx_n1.json 2> $O/bench_fnx_n1.err; echo "bench rc=$?"
python bench.py --frames-in-flight 1 --lanes 1 --no-cpu-baseline --no-dropin > $O/bench_fnx_n1_oneframe.json 2>/dev/null
export FNX_WORKLOAD=smoke
bash tools/gpu_ncu.sh s2final2/smoke merge_bucket_kernel density_bwd_kernel density_fwd_counted_kernel > /dev/null 2>&1
timeout 300 python tools/profile_step.py > $O/smoke/profile_step.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ls $O $O/smoke
