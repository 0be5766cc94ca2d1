#!/bin/bash
O=gpurun_out/s2a; mkdir -p $O
B="python bench.py --no-ab --no-dropin --no-cpu-baseline --min-leg-seconds 1.0"
for L in 4 6 8 16; do
  timeout 300 $B --lanes $L > $O/lanes$L.json 2> $O/lanes$L.err
  python - <<PY
import json
try:
    d=json.load(open("$O/lanes$L.json")); print("lanes $L value", d["value"], "e2e", d["e2e"]["value"], "lanes1", d.get("value_lanes1"), "lat", d.get("latency_one_frame_ms"))
except Exception as e: print("lanes $L failed", e)
PY
done
FNX_LPT_FWD=1 timeout 300 $B --lanes 4 > $O/lptfwd.json 2> $O/lptfwd.err
python - <<PY
import json
d=json.load(open("$O/lptfwd.json")); print("LPT_FWD value", d["value"], "e2e", d["e2e"]["value"], "lanes1", d.get("value_lanes1"), "lat", d.get("latency_one_frame_ms"))
PY
timeout 300 $B --lanes 8 --frames-in-flight 32 > $O/l8f32.json 2> $O/l8f32.err
python - <<PY
import json
d=json.load(open("$O/l8f32.json")); print("l8 f32 value", d["value"], "e2e", d["e2e"]["value"])
PY
