#!/bin/bash
# A/B of two libfnx builds on the GPU box: FNX_LIBFNX=<base> vs the in-tree build.  usage: exp_ab.sh <tag> [workloads...]
TAG=$1; shift
WLS=${@:-smoke}
O=gpurun_out/$TAG; mkdir -p $O
B="python bench.py --no-ab --no-dropin --no-cpu-baseline --min-leg-seconds 1.0"
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt; tail -3 $O/pytest.txt
fi
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d.get("roofline", {})
    print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "lanes1", d.get("value_lanes1"), "lat", d.get("latency_one_frame_ms"),
          "|", r.get("kernel"), r.get("ms_per_launch"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
for wl in $WLS; do
  if [ -f fluidnexus_b200/libfnx_base.so ]; then
    FNX_LIBFNX=$PWD/fluidnexus_b200/libfnx_base.so timeout 300 $B --workload $wl > $O/base_$wl.json 2> $O/base_$wl.err; show "base $wl" $O/base_$wl.json
  fi
  timeout 300 $B --workload $wl > $O/new_$wl.json 2> $O/new_$wl.err; show "new  $wl" $O/new_$wl.json
done
