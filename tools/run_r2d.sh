mkdir -p gpurun_out/r2d
for wl in smoke scalar c2 ball; do
  python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/r2d/bench_${wl}_n1.json 2> gpurun_out/r2d/bench_${wl}_n1.err
done
for wl in scalar ball c2; do
  python bench.py --impl reference --workload $wl --steps 10 --warmup 2 --ref-max-seconds 60 > gpurun_out/r2d/ref_${wl}.json 2> gpurun_out/r2d/ref_${wl}.err
done
python bench.py --lanes 8 --frames-in-flight 16 --no-cpu-baseline --no-dropin --no-ab --steps 20 > gpurun_out/r2d/bench_smoke_lanes8.json 2>/dev/null
python bench.py --lanes 4 --frames-in-flight 32 --no-cpu-baseline --no-dropin --no-ab --steps 10 > gpurun_out/r2d/bench_smoke_g32.json 2>/dev/null
ls -la gpurun_out/r2d
