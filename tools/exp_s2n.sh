#!/bin/bash
O=gpurun_out/s2n; mkdir -p $O
timeout 600 python -m pytest tests/test_bench_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-dropin --no-cpu-baseline > $O/bench_fnx_n2.json 2> $O/bench_fnx_n2.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench_fnx_n2.json")); print("N=2 value", d["value"], "e2e", d["e2e"]["value"], "cache", d["e2e"]["with_gt_cache"]["value"], "parity", d["sharding_parity"])
PY
