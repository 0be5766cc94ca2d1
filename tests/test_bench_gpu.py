"""GPU: bench.py end to end on the `tiny` workload -- the JSON contract of the line it prints, the sharded-vs-single-rank parity
assertion (VERDICT r1 "weak" #1: the flat gradient bucket must not accumulate across steps), and, when the box has >= 2 GPUs,
the same under torchrun with 2 NCCL ranks in both sharding modes (whole frames per rank; views of one frame split over ranks)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ["--workload", "tiny", "--steps", "4", "--warmup", "3", "--min-leg-seconds", "0.2", "--no-cpu-baseline", "--no-dropin", "--verify"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(n, extra):
    if n == 1:
        cmd = [sys.executable, "bench.py", "--gpus", "1", *COMMON, *extra]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), "bench.py", "--gpus", str(n), *COMMON, *extra]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-4000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def _check_line(d, n, G):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "value_lanes1", "leg", "sharding_parity"):
        assert k in d, k
    assert d["n_gpus"] == n and d["config"]["frames_in_flight"] == G and d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0
    assert d["leg"]["intervals"] >= 5 and d["e2e"]["leg"]["intervals"] >= 5
    sp = d["sharding_parity"]
    assert sp["ok"] is True and sp["max_abs_param_delta"] <= sp["tolerance"] and sp["loss_rel_delta_last_step"] < 1e-4
    r = d["roofline"]
    assert r["bound"] == "hbm" and 0 < r["frac"] < 1 and r["ms_per_launch"] > 0


def test_bench_tiny_single_gpu():
    d = _run(1, ["--frames-in-flight", "4", "--lanes", "2"])
    _check_line(d, 1, 4)
    assert d["latency_one_frame_ms"] > 0 and d["static_tile_cache"]["value_with_cache_off"] > 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("G", [4, 1])
def test_bench_tiny_two_nccl_ranks(G):
    """G = 4: two whole frames per rank (no gradient exchange, local Adam).  G = 1: the five views of one frame split 3 + 2 over
    the ranks, gradient all-reduce, replicated Adam.  Either way the parameters after 3 steps must equal a single-rank run's."""
    d = _run(2, ["--frames-in-flight", str(G), "--lanes", "2"])
    _check_line(d, 2, G)
    assert ("split over 2 ranks" in d["config"]["parallelism"]) == (G == 1)
