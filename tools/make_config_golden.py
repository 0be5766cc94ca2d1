"""Run in the build container (needs /root/reference): the EFFECTIVE optimisation constants of the reference's scene configs
(FluidDynamics/arguments/__init__.py defaults overridden by configs/*.json) for the keys this package hard-codes in
StepParams / PBFSolver / bench.py's workloads -> tests/golden/pyref_configs.json."""
import json
import os
import sys
from argparse import ArgumentParser

REF = "/root/reference/FluidDynamics"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pyref_configs.json")
KEYS = ["H", "KNN_K", "p0", "k", "secs", "alpha", "buoyancy_max_y", "buoyancy_decay_rate", "lambda_dssim", "lambda_image", "lambda_current_distance",
        "lambda_exyz", "lambda_gas_constraints", "lambda_next_gas_constraints", "distance_threshold_visual", "position_lr_init",
        "position_lr_final", "position_lr_delay_mult", "position_lr_max_steps", "color_lr", "opacity_lr", "scaling_lr", "rotation_lr",
        "percent_dense", "solver_iterations", "min_neighbors", "batch"]


def main():
    sys.path.insert(0, REF)
    import arguments
    defaults = arguments.OptimizationParams(ArgumentParser())
    out = {"defaults": {k: getattr(defaults, k) for k in KEYS}}
    for cfg in ("fluid_nexus_smoke_dynamics", "scalar_real", "fluid_nexus_ball_dynamics", "fluid_nexus_smoke_background"):
        c = json.load(open(os.path.join(REF, "configs", cfg + ".json")))
        out[cfg] = {k: (c[k] if k in c else getattr(defaults, k)) for k in KEYS}
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(out["fluid_nexus_smoke_dynamics"]))


if __name__ == "__main__":
    main()
