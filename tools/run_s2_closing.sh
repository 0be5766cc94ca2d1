#!/bin/bash
# round 2, closing build (matrix-form geometry backward, _C-level drop-in): GPU suite + headline bench line + smoke, inside the last GPU minutes
O=gpurun_out/s2final4
mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q --maxfail=6 2>&1 | tail -25 > $O/pytest_final.txt; tail -3 $O/pytest_final.txt
timeout 120 python bench.py --steps 20 --warmup 5 --no-dropin > $O/bench_fnx_n1.json 2> $O/bench_fnx_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/s2final4/bench_fnx_n1.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "lanes1", d.get("value_lanes1"), "lat", d.get("latency_one_frame_ms"), "sections", d.get("sections_ms_per_step"))
except Exception as e:
    print("bench line unreadable", e)
PY
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
