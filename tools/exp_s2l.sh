#!/bin/bash
O=gpurun_out/s2l; mkdir -p $O
B="python bench.py --no-ab --no-dropin --no-cpu-baseline --min-leg-seconds 1.0"
for wl in scalar c2; do for bb in 0 1; do
FNX_BUCKET_BINNING=$bb timeout 300 $B --workload $wl > $O/${wl}_bb$bb.json 2> $O/${wl}_bb$bb.err
python - <<PY
import json
try:
    d=json.load(open("$O/${wl}_bb$bb.json")); print("$wl bucket=$bb value", d["value"], "e2e", d["e2e"]["value"], "lanes1", d.get("value_lanes1"), "lat", d.get("latency_one_frame_ms"), {k:round(v,2) for k,v in d.get("sections_ms_per_step",{}).items()})
except Exception as e: print("$wl $bb failed", e)
PY
done; done
