"""Run in the build container (needs /root/reference): executes the reference's OWN particle-physics methods
(FluidDynamics/gaussian_splatting/gm_fluid.py: get_visual_xyz_from_nn, get_gas_constraints_from_exyz_nn,
get_guess_hidden_particles_from_nn, get_gas_constraints_from_vel_nn_guess, project_gas_constraints, update_visual_particles,
remove_invalid_particles) on the CPU and stores inputs + outputs in tests/golden/pyref_physics.npz.

gm_fluid.py imports torch_cluster / torch_scatter / simple_knn at module top; none is installable here.  They are replaced,
for this script only, by stub modules whose radius / radius_graph / scatter_min are the restatements in oracle/pbf_ref.py --
so the fixture pins EVERYTHING the reference computes around the neighbour search (kernels, index conventions, scatter sums,
the autograd chain, the solver update), while the neighbour search itself stays the documented unpinned part.  The model object
is created without its constructor (which needs a GPU) and given exactly the attributes the methods read."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/FluidDynamics"
OUT = os.path.join(ROOT, "tests", "golden", "pyref_physics.npz")
sys.path.insert(0, ROOT)
from fluidnexus_b200 import synthetic as S  # noqa: E402
from oracle import pbf_ref as O  # noqa: E402


def install_stubs():
    tc = types.ModuleType("torch_cluster")
    tc.radius = lambda x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, **kw: O.radius(x, y, r, max_num_neighbors=max_num_neighbors)
    tc.radius_graph = lambda x, r, batch=None, loop=False, max_num_neighbors=32, **kw: O.radius_graph(x, r, loop=loop, max_num_neighbors=max_num_neighbors)
    ts = types.ModuleType("torch_scatter")
    ts.scatter_min = lambda src, index, dim=0, dim_size=None: O.scatter_min(src, index, dim=dim, dim_size=dim_size)
    sk, skc = types.ModuleType("simple_knn"), types.ModuleType("simple_knn._C")
    skc.distCUDA2 = lambda pts: torch.tensor(O.knn3_mean_dist2(pts.numpy()))
    sk._C = skc
    sys.modules.update({"torch_cluster": tc, "torch_scatter": ts, "simple_knn": sk, "simple_knn._C": skc})


import contextlib


@contextlib.contextmanager
def cuda_as_cpu():
    """The reference hard-wires device="cuda" / .cuda() in some methods; inside this context those land on the CPU."""
    names = ["zeros", "ones", "tensor", "empty", "full", "arange", "zeros_like", "ones_like", "normal", "rand", "randn"]
    saved = {n: getattr(torch, n) for n in names}

    def wrap(fn):
        def inner(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return fn(*a, **k)
        return inner
    saved_cuda = torch.Tensor.cuda
    try:
        for n in names:
            setattr(torch, n, wrap(saved[n]))
        torch.Tensor.cuda = lambda self, *a, **k: self
        yield
    finally:
        for n in names:
            setattr(torch, n, saved[n])
        torch.Tensor.cuda = saved_cuda


def make_model(GM, K, p0, bmax, seed):
    hp = S.hidden_lattice(600, seed=seed, buoyancy=(0.0, 1.96, 0.0))
    rng = np.random.default_rng(seed)
    f = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    gm = object.__new__(GM)
    gm.H, gm.KNN_K, gm.p0, gm.k, gm._secs, gm.scale_factor, gm.buoyancy_max_y = 2.0, K, p0, 3.0, 0.033, 100.0, bmax
    gm.H2, gm.H6, gm.H9 = gm.H ** 2, gm.H ** 6, gm.H ** 9
    gm.EPSILON, gm.RELAXATION, gm.K_P, gm.E_P, gm.DQ_P = 1e-8, 0.01, 0.2, 4, 0.25
    gm.poly6_term1 = 315.0 / (64.0 * np.pi * gm.H9)
    gm.spiky_grad_term1 = 45.0 / (np.pi * gm.H6)
    gm.lamb_corr_denom = gm.poly6(torch.tensor(gm.DQ_P * gm.DQ_P * gm.H * gm.H, dtype=torch.float64))
    gm.record_time, gm.min_neighbors = False, 20
    gm._xyz, gm._estimate_xyz = f(hp.xyz), f(hp.estimate_xyz)
    gm._buoyancy, gm._force = f(hp.buoyancy), f(rng.normal(0, 3, hp.force.shape))
    gm._velocity = f(rng.normal(0, 5, hp.xyz.shape) + np.array([0.0, 30.0, 0.0]))
    gm._imass = f(rng.uniform(0.8, 1.2, (hp.N, 1)))
    gm._counts = torch.full((hp.N, 1), 2.0, dtype=torch.float64)
    gm._particle_id = torch.arange(hp.N)
    vis = hp.xyz[rng.choice(hp.N, 150, replace=False)] + rng.uniform(-0.4, 0.4, (150, 3))
    gm._visual_xyz = f(vis)
    gm._estimate_xyz_nn = (gm._estimate_xyz / gm.scale_factor).clone().requires_grad_(True)
    return gm


def main():
    install_stubs()
    # the reference allocates its scatter targets with the default dtype: run its code in fp64 end to end
    torch.set_default_dtype(torch.float64)
    sys.path.insert(0, REF)
    from gaussian_splatting.gm_fluid import GaussianModel as GM
    out = {}
    for tag, (K, p0, bmax) in {"smoke": (100, 1.5, 0.0), "scalar": (100, 2.0, 0.8), "capped": (14, 1.5, 0.8)}.items():
        gm = make_model(GM, K, p0, bmax, seed=5)
        state0 = {k: getattr(gm, "_" + k).detach().numpy().copy() for k in ("xyz", "estimate_xyz", "buoyancy", "force", "velocity", "imass", "counts",
                                                                             "visual_xyz")}
        # ---- the differentiable terms of the training step and their gradient w.r.t. estimate_xyz_nn ----
        rng = np.random.default_rng(1)
        w_vis = torch.tensor(rng.normal(size=(150, 3)))
        vis_out = gm.get_visual_xyz_from_nn()
        p2 = gm.get_gas_constraints_from_exyz_nn()
        Y = gm.get_guess_hidden_particles_from_nn()
        p3 = gm.get_gas_constraints_from_vel_nn_guess()
        loss = (vis_out * w_vis).sum() + ((p2 - 1.0) ** 2).mean() + 0.1 * ((p3 - 1.0) ** 2).mean()
        loss.backward()
        out.update({f"{tag}_{k}": v for k, v in state0.items()})
        out.update({f"{tag}_K": K, f"{tag}_p0": p0, f"{tag}_bmax": bmax, f"{tag}_w_vis": w_vis.numpy(), f"{tag}_P1": vis_out.detach().numpy(),
                    f"{tag}_P2": p2.detach().numpy(), f"{tag}_Y": Y.detach().numpy(), f"{tag}_P3": p3.detach().numpy(),
                    f"{tag}_loss": loss.item(), f"{tag}_grad": gm._estimate_xyz_nn.grad.numpy().copy()})
        # ---- no-grad solver pieces (in place on the model) ----
        ret = gm.project_gas_constraints()
        out.update({f"{tag}_proj_estimate_xyz": gm._estimate_xyz.numpy().copy(), f"{tag}_proj_force": gm._force.numpy().copy(),
                    f"{tag}_proj_lambda_mean": ret["lambdas"], f"{tag}_proj_pratio_mean": ret["p_ratio"]})
        gm.update_visual_particles()
        out[f"{tag}_visual_after_update"] = gm._visual_xyz.numpy().copy()
        n0 = gm._xyz.shape[0]
        gm.remove_invalid_particles()
        out[f"{tag}_kept_after_prune"] = np.int64(gm._xyz.shape[0])
        out[f"{tag}_n0"] = np.int64(n0)
    # ---- the FluidNexus scenes use gm_dynamics.GaussianModel (same physics code, gm_dynamics.py:1014-1030,1269-1320,1453-1498):
    #      its methods must give what gm_fluid's gave ----
    ply = types.ModuleType("plyfile")
    ply.PlyData, ply.PlyElement = object, object
    sys.modules["plyfile"] = ply
    from gaussian_splatting.gm_dynamics import GaussianModel as GD
    same = True
    for tag, (K, p0, bmax) in {"smoke": (100, 1.5, 0.0), "scalar": (100, 2.0, 0.8), "capped": (14, 1.5, 0.8)}.items():
        gd = make_model(GD, K, p0, bmax, seed=5)
        same &= bool(np.array_equal(gd.get_visual_xyz_from_nn().detach().numpy(), out[f"{tag}_P1"]))
        same &= bool(np.array_equal(gd.get_gas_constraints_from_exyz_nn().detach().numpy(), out[f"{tag}_P2"]))
        same &= bool(np.array_equal(gd.get_gas_constraints_from_vel_nn_guess().detach().numpy(), out[f"{tag}_P3"]))
    assert same, "gm_dynamics and gm_fluid disagree on P1-P3"
    out["gm_dynamics_bit_identical"] = np.bool_(same)
    # ---- guess_hidden_particles / confirm_guess_hidden_particles (device="cuda" redirected to the CPU) ----
    for tag, (bmax, stable, wind, decay) in {"guess_plain": (0.0, False, False, 0.0), "guess_bmax_wind": (0.8, False, True, 0.9),
                                              "guess_stable": (0.8, True, False, 0.0)}.items():
        torch.set_default_dtype(torch.float32)   # these two methods hard-wire torch.float: run them in the reference's own precision
        gm = make_model(GM, 100, 1.5, bmax, seed=9)
        for k in ("xyz", "estimate_xyz", "velocity", "force", "buoyancy", "imass", "counts", "visual_xyz"):
            setattr(gm, "_" + k, getattr(gm, "_" + k).float())
        gm._gravity = torch.tensor([0.0, -9.8, 0.0]).reshape((1, 3))
        gm.alpha, gm.buoyancy_decay_rate = -0.2, decay
        gm.wind_force, gm.wind_force_max, gm.wind_power = torch.tensor([0.3, 0.0, 0.1]).reshape((1, 3)), 0.3, 2.0
        out.update({f"{tag}_{k}": getattr(gm, "_" + k).detach().numpy().copy() for k in ("xyz", "velocity", "force", "buoyancy")})
        with cuda_as_cpu():
            gm.guess_hidden_particles(stable=stable, use_wind=wind)
            after_guess = {k: getattr(gm, "_" + k).detach().numpy().astype(np.float64).copy() for k in ("velocity", "force", "buoyancy", "estimate_xyz", "counts")}
            # a solver-like displacement, with every third particle left exactly where it was
            gm._estimate_xyz = gm._estimate_xyz + 0.05 * torch.sin(gm._xyz)
            gm._estimate_xyz[::3] = gm._xyz[::3]
            moved = gm._estimate_xyz.numpy().copy()
            gm.confirm_guess_hidden_particles()
        out.update({f"{tag}_after_{k}": v for k, v in after_guess.items()})
        out.update({f"{tag}_moved": moved, f"{tag}_confirm_xyz": gm._xyz.numpy().copy(), f"{tag}_confirm_velocity": gm._velocity.numpy().copy(),
                    f"{tag}_bmax": bmax, f"{tag}_stable": stable, f"{tag}_wind": wind, f"{tag}_decay": decay})
    # ---- P6 / P7: gradient cache over the view loop, 1/batch, Adam as training_setup_current configures it ----
    torch.set_default_dtype(torch.float32)
    gm = make_model(GM, 100, 1.5, 0.0, seed=13)
    gm._estimate_xyz = gm._estimate_xyz.float()
    gm.spatial_lr_scale = 1.0

    class OptimArgs:
        position_lr_init, position_lr_final, position_lr_delay_mult, position_lr_max_steps = 1.6e-4, 1.6e-6, 0.01, 30_000
    rng = np.random.default_rng(21)
    with cuda_as_cpu():
        gm.training_setup_current(OptimArgs)
        out["plumb_e0"] = gm._estimate_xyz_nn.detach().numpy().copy()
        out["plumb_lr"] = gm.optimizer.param_groups[0]["lr"]
        out["plumb_eps"] = gm.optimizer.param_groups[0]["eps"]
        view_grads = rng.normal(0, 1e-2, (2, 3) + tuple(gm._estimate_xyz_nn.shape)).astype(np.float32)     # 2 iterations x 3 views
        for it in range(2):
            gm.update_learning_rate_current(it + 1)                # computes a rate but never assigns it (gm_fluid.py:401-407)
            gm.zero_gradient_cache_current()
            for v in range(3):
                gm._estimate_xyz_nn.grad = torch.tensor(view_grads[it, v])
                gm.cache_gradient_current()
                gm.optimizer.zero_grad(set_to_none=True)
            gm.set_batch_gradient_current(3)
            out[f"plumb_batch_grad{it}"] = gm._estimate_xyz_nn.grad.numpy().copy()
            gm.optimizer.step()
            out[f"plumb_e{it + 1}"] = gm._estimate_xyz_nn.detach().numpy().copy()
        out["plumb_lr_after"] = gm.optimizer.param_groups[0]["lr"]
    out["plumb_view_grads"] = view_grads
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "with", len(out), "arrays;", {k: int(out[k]) for k in out if k.endswith("kept_after_prune")})


if __name__ == "__main__":
    main()
