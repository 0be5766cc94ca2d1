"""Dev tool: time of one no-grad PBF simulation tick (fluidnexus_b200.solver) at the bench sizes (N = 28k hidden, 20k visual)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluidnexus_b200 import synthetic as S  # noqa: E402
from fluidnexus_b200.solver import PBFSolver  # noqa: E402


def main():
    hp = S.hidden_lattice(28_000, seed=1)
    rng = np.random.default_rng(0)
    vel = rng.normal(0, 5, hp.xyz.shape) + np.array([0.0, 30.0, 0.0])
    vis = hp.xyz[rng.choice(hp.N, 20_000, replace=False)] + rng.uniform(-0.4, 0.4, (20_000, 3))
    sol = PBFSolver(hp.xyz, velocity=vel, visual_xyz=vis)
    for iters in (3, 10):
        for _ in range(3):
            sol.tick(solver_iterations=iters)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            sol.tick(solver_iterations=iters)
        e1.record()
        torch.cuda.synchronize()
        print(f"PBF tick, N={hp.N}, V=20000, {iters} solver iterations: {e0.elapsed_time(e1) / n:.3f} ms per tick "
              f"({e0.elapsed_time(e1) / n / iters:.3f} ms per solver iteration incl. guess/confirm/visual update)")


if __name__ == "__main__":
    main()
