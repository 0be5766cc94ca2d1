"""oracle/pbf_ref.py -- torch-CPU restatement of the particle-physics terms.  TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline/reference legs and __graft_entry__.smoke() may import this; the product
(fluidnexus_b200/) never does.

Restates, op for op, from the reference (FD = FluidDynamics):
  poly6 kernel                               FD/gaussian_splatting/gm_fluid.py:124,166-169
  P1 get_visual_xyz_from_nn                  gm_fluid.py:1291-1336
  P2 get_gas_constraints_from_exyz_nn        gm_fluid.py:1107-1132
  P3 get_gas_constraints_from_vel_nn_guess   gm_fluid.py:1134-1158 + get_guess_hidden_particles_from_nn :846-862
  P4 exyz tie, step loss assembly            FD/entries_scalar_real/train_physical_particle.py:328-355
  P5 distance_loss                           FD/utils/loss_utils.py:98-121
  P6 grad cache / batch average              gm_fluid.py:409-430
  P7 Adam(eps=1e-15), lr never updated       gm_fluid.py:330-355,401-407

PINNED by the reference's own code, for everything except the neighbour search: tools/make_physics_golden.py and
tools/make_python_golden.py import the reference's modules from /root/reference on the CPU (gm_fluid.py with torch_cluster /
torch_scatter / simple_knn replaced by stubs that call `radius` / `radius_graph` / `scatter_min` below; utils/loss_utils.py
as it is) and store what ITS methods compute -- P1-P3 with their gradient, project_gas_constraints, update_visual_particles,
remove_invalid_particles, guess / confirm, the gradient cache and Adam set-up, l1 / ssim / distance_loss -- in
tests/golden/pyref_physics.npz and pyref_python.npz; tests/test_reference_physics_golden.py and
tests/test_reference_python_golden.py hold this file to those numbers (1e-12 in fp64).

PARITY UNPINNED for the neighbour search itself: the reference calls `torch_cluster.radius / radius_graph`
(torch-cluster==1.6.3, fluid_nexus.yml:397) and `torch_scatter.scatter_min` (torch-scatter==2.1.2, :399), third-party
CUDA extensions that are neither vendored in /root/reference nor installed here, and the reference has no tests
or golden vectors for this path.  `radius` / `radius_graph` below restate the published
algorithm of torch-cluster 1.6.3's CUDA kernel (csrc/cuda/radius_cuda.cu): for each query y_c scan x in index
order, accept j iff sum_d (x_j - y_c)^2 < r^2 (strict), stop after max_num_neighbors hits; radius_graph(x, r, loop,
K) = radius(x, x, r, K if loop else K+1), rows swapped for flow='source_to_target' (row = neighbour, col = query),
self pairs dropped when loop=False.  Anchors are the reference's own call sites listed above and closed-form
checks (two-particle density = poly6(0) + poly6(d^2), ...) in tests/test_oracle_pbf.py.

`distance_loss` is restated with exact coordinate differences; the reference's `torch.cdist` switches to the
|a|^2+|b|^2-2ab matmul form for more than 25 points, which carries ~1e-3 relative noise at the distances this
term looks at (and needs O(V^2) memory).
"""
import math

import numpy as np
import torch

SCALE_FACTOR = 100.0  # gm_fluid.py:122
EPSILON = 1e-8        # gm_fluid.py:99


# ------------------------------------------------------------------------------------------------------------
# torch_cluster / torch_scatter semantics
# ------------------------------------------------------------------------------------------------------------
def radius(x, y, r, max_num_neighbors=32):
    """edge_index[0] = index into y (query), [1] = index into x; first `max_num_neighbors` hits in x-index order."""
    xn = x.detach().cpu().numpy().astype(np.float32)
    yn = y.detach().cpu().numpy().astype(np.float32)
    r2 = np.float32(r) * np.float32(r)
    if xn.shape[0] == 0 or yn.shape[0] == 0:
        return torch.zeros((2, 0), dtype=torch.long)
    if xn.shape[0] * yn.shape[0] <= 250_000:  # literal scan, the form of torch-cluster's CUDA kernel
        rows, cols = [], []
        for c in range(yn.shape[0]):
            d = ((xn - yn[c]) ** 2).sum(1, dtype=np.float32)
            hit = np.nonzero(d < r2)[0][:max_num_neighbors]
            rows.append(np.full(hit.shape, c, np.int64))
            cols.append(hit.astype(np.int64))
        return torch.from_numpy(np.stack([np.concatenate(rows), np.concatenate(cols)]))
    # same result through a k-d tree candidate search (vectorised): candidates within r(1+eps), exact fp32 test after
    from scipy.spatial import cKDTree
    tx, ty = cKDTree(xn.astype(np.float64)), cKDTree(yn.astype(np.float64))
    pairs = ty.sparse_distance_matrix(tx, float(r) * (1 + 1e-5) + 1e-6, output_type="ndarray")
    q, j = pairs["i"].astype(np.int64), pairs["j"].astype(np.int64)
    d = ((xn[j] - yn[q]) ** 2).sum(1, dtype=np.float32)
    keep = d < r2
    q, j = q[keep], j[keep]
    order = np.lexsort((j, q))
    q, j = q[order], j[order]
    if q.size:
        start = np.r_[0, np.nonzero(np.diff(q))[0] + 1]
        rank = np.arange(q.size) - np.repeat(start, np.diff(np.r_[start, q.size]))
        keep = rank < max_num_neighbors
        q, j = q[keep], j[keep]
    return torch.from_numpy(np.stack([q, j]))


def radius_graph(x, r, loop=False, max_num_neighbors=32):
    e = radius(x, x, r, max_num_neighbors if loop else max_num_neighbors + 1)
    row, col = e[1], e[0]  # flow == 'source_to_target'
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col])


def scatter_min(src, index, dim=0, dim_size=None):
    """torch_scatter.scatter_min for 1-D src: (min per index, argmin; empty groups -> (0, src.numel()))."""
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros(n, dtype=src.dtype)
    arg = torch.full((n,), src.numel(), dtype=torch.long)
    s, ix = src.detach().numpy(), index.numpy()
    best = {}
    for k in range(s.shape[0]):
        i = int(ix[k])
        if i not in best or s[k] < s[best[i]]:
            best[i] = k
    for i, k in best.items():
        out[i] = float(s[k])
        arg[i] = k
    return out, arg


# ------------------------------------------------------------------------------------------------------------
# physics terms
# ------------------------------------------------------------------------------------------------------------
class PBFParams:
    """The constants the reference reads from `optim_args` (FD/arguments/__init__.py:308,312 + configs)."""

    def __init__(self, H=2.0, KNN_K=100, p0=1.5, secs=0.033, buoyancy_max_y=0.0, lambda_dssim=0.2, lambda_image=1.0,
                 lambda_current_distance=0.1, lambda_exyz=0.1, lambda_gas_constraints=1.0,
                 lambda_next_gas_constraints=0.1, distance_threshold_visual=0.002, lr=1.6e-4):
        self.H, self.KNN_K, self.p0, self.secs, self.buoyancy_max_y = H, KNN_K, p0, secs, buoyancy_max_y
        self.lambda_dssim, self.lambda_image = lambda_dssim, lambda_image
        self.lambda_current_distance, self.lambda_exyz = lambda_current_distance, lambda_exyz
        self.lambda_gas_constraints, self.lambda_next_gas_constraints = lambda_gas_constraints, lambda_next_gas_constraints
        self.distance_threshold_visual, self.lr = distance_threshold_visual, lr
        self.H2 = H * H
        self.poly6_term1 = 315.0 / (64.0 * np.pi * H ** 9)  # gm_fluid.py:124


def poly6(prm, r2):
    """gm_fluid.py:166-169.  (The mask is cast to r2's dtype: `bool * python float` would round the constant to the default
    dtype, fp32, inside an fp64 evaluation.)"""
    term2 = prm.H2 - r2
    mask = (r2 < prm.H2).to(r2.dtype)
    return mask * prm.poly6_term1 * (term2 ** 3)


def visual_xyz_from_nn(prm, estimate_xyz_nn, xyz, visual_xyz):
    """P1.  estimate_xyz_nn [N,3] (trainable, render units), xyz [N,3], visual_xyz [V,3] (scaled units)."""
    visual_xyz = visual_xyz.detach()
    est = estimate_xyz_nn * SCALE_FACTOR
    velocity_nn = (est - xyz) / prm.secs
    V = visual_xyz.shape[0]
    e = radius(x=est, y=visual_xyz, r=prm.H, max_num_neighbors=prm.KNN_K)
    row, col = e[0], e[1]
    diff = visual_xyz[row] - est[col]
    dist2 = torch.sum(diff ** 2, dim=1)
    p6 = poly6(prm, dist2)
    weighted_velocity = velocity_nn[col] * p6.unsqueeze(-1)
    visual_velocity = torch.zeros(V, 3, dtype=est.dtype).index_add_(0, row, weighted_velocity)
    sum_p6 = torch.zeros(V, dtype=est.dtype).index_add_(0, row, p6).clamp_min(EPSILON)
    return visual_xyz + visual_velocity * prm.secs / sum_p6.unsqueeze(-1)


def _density_ratio(prm, pos, imass):
    N = pos.shape[0]
    e = radius_graph(pos, r=prm.H, loop=True, max_num_neighbors=prm.KNN_K)
    row, col = e
    diff = pos[row] - pos[col]
    dist2 = torch.sum(diff ** 2, dim=1)
    pi = torch.zeros(N, dtype=pos.dtype).index_add_(0, row, poly6(prm, dist2))
    return pi.unsqueeze(1) / imass / prm.p0


def gas_constraints_from_exyz_nn(prm, estimate_xyz_nn, imass):
    """P2: p_ratio [N,1] at the optimised positions."""
    return _density_ratio(prm, estimate_xyz_nn * SCALE_FACTOR, imass)


def guess_hidden_particles_from_nn(prm, estimate_xyz_nn, xyz, buoyancy, force):
    """gm_fluid.py:846-862 (note: the UNSCALED parameter's y drives the buoyancy coefficient)."""
    if prm.buoyancy_max_y > 0.0:
        cur_buoyancy = buoyancy * (1.0 - (estimate_xyz_nn[:, 1:2] / prm.buoyancy_max_y))
    else:
        cur_buoyancy = buoyancy
    tmp_velocity = (estimate_xyz_nn * SCALE_FACTOR - xyz) / prm.secs
    estimate_velocity = tmp_velocity + cur_buoyancy * prm.secs + prm.secs * force
    return estimate_xyz_nn * SCALE_FACTOR + prm.secs * estimate_velocity


def gas_constraints_from_vel_nn_guess(prm, estimate_xyz_nn, xyz, buoyancy, force, imass):
    """P3: p_ratio [N,1] one advection tick later."""
    return _density_ratio(prm, guess_hidden_particles_from_nn(prm, estimate_xyz_nn, xyz, buoyancy, force), imass)


def l2_loss(a, b):
    return ((a - b) ** 2).mean()


def distance_loss(positions, threshold):
    """P5 (loss_utils.py:98-121) with exact differences instead of cdist's matmul form; O(V^2) memory."""
    d = positions.unsqueeze(1) - positions.unsqueeze(0)
    d2 = (d ** 2).sum(-1)
    eye = torch.eye(positions.shape[0], dtype=torch.bool)
    dist = torch.sqrt(torch.where(d2 > 0, d2, torch.ones_like(d2)))  # keep sqrt'(0) = inf out of the graph
    mask = (dist < threshold) & ~eye & (d2 > 0)
    # exact duplicates (d == 0): value (thr-0)^2 with zero gradient, as cdist's backward gives
    dup = (~eye) & (d2 <= 0)
    loss = (((threshold - dist) * mask.to(dist.dtype)).clamp(min=0) ** 2).sum()
    return loss + dup.sum().to(dist.dtype) * threshold ** 2


def physics_loss_terms(prm, estimate_xyz_nn, st, with_distance=True, visual_xyz=None):
    """The view-independent part of the step loss (train_physical_particle.py:331-355) and the advected visual
    positions.  `st` holds torch tensors xyz, estimate_xyz, buoyancy, force, imass [,visual_xyz]."""
    vis = visual_xyz_from_nn(prm, estimate_xyz_nn, st["xyz"], st["visual_xyz"] if visual_xyz is None else visual_xyz)
    render_xyz = vis / SCALE_FACTOR
    terms = {}
    terms["dist"] = distance_loss(render_xyz, prm.distance_threshold_visual) if with_distance else torch.zeros(())
    terms["exyz"] = l2_loss(estimate_xyz_nn * SCALE_FACTOR, st["estimate_xyz"])
    p_ratio = gas_constraints_from_exyz_nn(prm, estimate_xyz_nn, st["imass"])
    terms["gas"] = l2_loss(p_ratio, torch.ones_like(p_ratio))
    p_next = gas_constraints_from_vel_nn_guess(prm, estimate_xyz_nn, st["xyz"], st["buoyancy"], st["force"], st["imass"])
    terms["next_gas"] = l2_loss(p_next, torch.ones_like(p_next))
    total = (prm.lambda_current_distance * terms["dist"] + prm.lambda_exyz * terms["exyz"]
             + prm.lambda_gas_constraints * terms["gas"] + prm.lambda_next_gas_constraints * terms["next_gas"])
    return total, terms, render_xyz, p_ratio, p_next


# ------------------------------------------------------------------------------------------------------------
# image losses (loss_utils.py:9-64) -- plain torch, used as the fp32 reference of the fused CUDA loss kernels
# ------------------------------------------------------------------------------------------------------------
def l1_loss(a, b):
    return torch.abs(a - b).mean()


def _window(window_size, channel):
    g = torch.tensor([math.exp(-((x - window_size // 2) ** 2) / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=11):
    import torch.nn.functional as F
    channel = img1.size(-3)
    window = _window(window_size, channel).to(img1.dtype).to(img1.device)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def image_loss(prm, image, gt, grey=False):
    """0.8*l1 + 0.2*(1-ssim) (train_physical_particle.py:328-329,346-347); `grey` applies the FluidNexus entries'
    channel-mean conversion to both images first (entries_fluid_nexus/train_physical_particle.py:356-360)."""
    if grey:
        gt = torch.cat([torch.mean(gt, dim=0, keepdim=True)] * 3, dim=0)
        image = torch.cat([torch.mean(image, dim=0, keepdim=True)] * 3, dim=0)
    l1 = l1_loss(image, gt)
    s = 1.0 - ssim(image, gt)
    return (1.0 - prm.lambda_dssim) * l1 * prm.lambda_image + prm.lambda_dssim * s * prm.lambda_image, l1, s


def knn3_mean_dist2(points):
    """distCUDA2 (KNN/simple_knn.cu:134-166): mean of the 3 smallest squared distances to OTHER points."""
    from scipy.spatial import cKDTree
    p = np.asarray(points, np.float64)
    d, _ = cKDTree(p).query(p, k=4)
    return (d[:, 1:4] ** 2).mean(1)


# ---------------------------------------------------------------------------------------------------------------------
# No-grad PBF solver tick.  Literal restatements of the reference methods on a plain state
# dict {xyz, estimate_xyz, velocity, force, buoyancy, imass [N,1], counts [N,1], visual_xyz}; every function mutates
# the dict the way the reference mutates `self`.  TEST INFRASTRUCTURE ONLY.  Parity unpinned at radius_graph (see the
# module header); everything else follows the reference line by line.
# ---------------------------------------------------------------------------------------------------------------------
class SolverParams:
    """gm_fluid.py:77-126 + FD/arguments/__init__.py:300-330."""

    def __init__(self, H=2.0, p0=1.5, k=10.0, KNN_K=100, secs=0.033, alpha=-0.2, buoyancy_max_y=0.0, buoyancy_decay_rate=0.0,
                 scale_factor=100.0, gravity=(0.0, -9.8, 0.0), wind_force=(0.0, 0.0, 0.0), wind_power=1.0, min_neighbors=-1):
        self.H, self.p0, self.k, self.KNN_K, self.secs, self.alpha = H, p0, k, KNN_K, secs, alpha
        self.buoyancy_max_y, self.buoyancy_decay_rate, self.scale_factor = buoyancy_max_y, buoyancy_decay_rate, scale_factor
        self.gravity, self.wind_force, self.wind_power = gravity, wind_force, wind_power
        self.wind_force_max = max(wind_force)
        self.min_neighbors = min_neighbors
        self.EPSILON, self.RELAXATION, self.K_P, self.E_P, self.DQ_P = 1e-8, 0.01, 0.2, 4, 0.25
        self.H2, self.H6, self.H9 = H ** 2, H ** 6, H ** 9
        self.poly6_term1 = 315.0 / (64.0 * np.pi * self.H9)
        self.spiky_grad_term1 = 45.0 / (np.pi * self.H6)
        t = self.H2 - self.DQ_P * self.DQ_P * H * H
        self.lamb_corr_denom = self.poly6_term1 * t ** 3   # poly6(DQ_P^2 H^2), gm_fluid.py:126


def spiky_grad(sp, r, rlen):
    """gm_fluid.py:171-177."""
    mask = (rlen < sp.H) & (rlen > 0)
    r_norm = r / (rlen.unsqueeze(-1) + sp.EPSILON)
    term2 = (sp.H - rlen).unsqueeze(-1) ** 2
    grad = -r_norm * sp.spiky_grad_term1 * term2
    grad[~mask] = 0.0
    return grad


def solver_guess_hidden_particles(sp, st, stable=False, use_wind=False):
    """gm_fluid.py:809-844."""
    cur_secs, cur_alpha = (0.01, -1.0) if stable else (sp.secs, sp.alpha)
    dt = st["xyz"].dtype
    g = torch.tensor(sp.gravity, dtype=dt).reshape(1, 3)
    st["buoyancy"] = torch.ones_like(st["buoyancy"]) * g * cur_alpha
    if sp.buoyancy_max_y > 0.0:
        coeff = 1.0 - (st["xyz"][:, 1:2] / (sp.buoyancy_max_y * sp.scale_factor))
        cur_buoyancy = st["buoyancy"] * coeff
    else:
        cur_buoyancy = st["buoyancy"]
    st["velocity"] = st["velocity"] + cur_buoyancy * cur_secs + cur_secs * st["force"]
    if use_wind:
        y_scale = st["xyz"][:, 1:2] / sp.scale_factor
        wf = (y_scale ** sp.wind_power) * torch.tensor(sp.wind_force, dtype=dt).reshape(1, 3)
        st["velocity"] = st["velocity"] + torch.clamp(wf, 0.0, sp.wind_force_max) * cur_secs
    if sp.buoyancy_decay_rate > 0.0:
        st["buoyancy"] = st["buoyancy"] * sp.buoyancy_decay_rate
    st["force"] = torch.zeros_like(st["force"])
    st["estimate_xyz"] = st["xyz"] + cur_secs * st["velocity"]
    st["counts"] = torch.zeros_like(st["counts"])


def solver_project_gas_constraints(sp, st):
    """gm_fluid.py:896-996 (the log dictionary of :998-1019 is not restated).  Returns (p_ratio, lambdas)."""
    exyz = st["estimate_xyz"]
    N = exyz.shape[0]
    row, col = radius_graph(exyz, sp.H, loop=True, max_num_neighbors=sp.KNN_K)
    non_self = row != col
    diff = exyz[row] - exyz[col]
    dist2 = torch.sum(diff ** 2, dim=1)
    mask = (dist2 < sp.H2).to(dist2.dtype)   # (keeps the fp64 evaluation fp64: bool * python float would round to fp32)
    poly6_values = mask * sp.poly6_term1 * ((sp.H2 - dist2) ** 3)
    pi = torch.zeros(N, dtype=exyz.dtype).index_add_(0, row, poly6_values).unsqueeze(1) / st["imass"]
    neighbors_len = torch.bincount(row, minlength=N).unsqueeze(1).to(exyz.dtype)
    row_ns, col_ns, diff_ns, dist2_ns = row[non_self], col[non_self], diff[non_self], dist2[non_self]
    rlen_ns = torch.sqrt(dist2_ns + sp.EPSILON)
    sg = spiky_grad(sp, diff_ns, rlen_ns)
    gr = torch.zeros(N, 3, dtype=exyz.dtype).index_add_(0, row_ns, sg) / sp.p0
    gr_dot = torch.sum(gr ** 2, dim=1)
    grad_dot = torch.zeros(N, dtype=exyz.dtype).index_add_(0, row_ns, torch.sum((sg / sp.p0) ** 2, dim=1))
    denom = (grad_dot + gr_dot).unsqueeze(1)
    p_ratio = pi / sp.p0
    st["force"] = st["force"] + st["velocity"] * (1.0 - p_ratio) * -sp.k
    lambdas = -(p_ratio - 1.0) / (denom + sp.RELAXATION)
    lamb_corr = -sp.K_P * (poly6_values[non_self] / sp.lamb_corr_denom) ** sp.E_P
    lam_sum = lambdas[row_ns].squeeze(1) + lambdas[col_ns].squeeze(1)
    deltas = (lam_sum + lamb_corr).unsqueeze(-1) * sg
    deltas_sum = torch.zeros(N, 3, dtype=exyz.dtype).index_add_(0, row_ns, deltas) / sp.p0
    st["estimate_xyz"] = exyz + deltas_sum / (neighbors_len + st["counts"])
    return p_ratio, lambdas


def solver_confirm_guess_hidden_particles(sp, st):
    """gm_fluid.py:1160-1175."""
    st["velocity"] = (st["estimate_xyz"] - st["xyz"]) / sp.secs
    mask = torch.norm(st["estimate_xyz"] - st["xyz"], dim=1) < sp.EPSILON
    st["velocity"][mask] = 0.0
    st["xyz"] = st["xyz"].clone()
    st["xyz"][~mask] = st["estimate_xyz"][~mask]


def solver_update_visual_particles(sp, st):
    """gm_fluid.py:1197-1239."""
    vis = st["visual_xyz"]
    if vis.shape[0] == 0:
        return
    e = radius(st["estimate_xyz"], vis, sp.H, max_num_neighbors=sp.KNN_K)
    row, col = e[0], e[1]
    diff = vis[row] - st["estimate_xyz"][col]
    dist2 = torch.sum(diff ** 2, dim=1)
    p6 = (dist2 < sp.H2).to(dist2.dtype) * sp.poly6_term1 * ((sp.H2 - dist2) ** 3)
    vv = torch.zeros(vis.shape[0], 3, dtype=vis.dtype).index_add_(0, row, st["velocity"][col] * p6.unsqueeze(-1))
    s = torch.zeros(vis.shape[0], dtype=vis.dtype).index_add_(0, row, p6).clamp_min(sp.EPSILON)
    st["visual_xyz"] = vis + vv * sp.secs / s.unsqueeze(-1)


def solver_neighbor_degree(sp, xyz):
    """bincount(row) of radius_graph(xyz, H, loop=False) with torch_cluster's default cap of 32 (gm_fluid.py:873-878)."""
    row, _ = radius_graph(xyz, sp.H, loop=False, max_num_neighbors=32)
    return torch.bincount(row, minlength=xyz.shape[0])


def solver_tick(sp, st, solver_iterations=3, stable=False, use_wind=False, count_first=False):
    """future_simulation.py:135-162 / train_physical_particle.py:206-216."""
    solver_guess_hidden_particles(sp, st, stable=stable, use_wind=use_wind)
    if count_first:
        st["counts"] = st["counts"] + float(solver_iterations)
    for _ in range(solver_iterations):
        solver_project_gas_constraints(sp, st)
    solver_confirm_guess_hidden_particles(sp, st)
    solver_update_visual_particles(sp, st)


def check_inside_rigid_body(kind, center, xyz, cuboid_num=None, particle_diameter=None, sphere_radius=None, cylinder_radius=None,
                            cylinder_num=None):
    """gm_fluid.py:1024-1056."""
    center = torch.as_tensor(center, dtype=xyz.dtype)
    if kind == "cuboid":
        half = torch.tensor([n * particle_diameter for n in cuboid_num], dtype=xyz.dtype) / 2.0
        return torch.all((xyz >= center - half) & (xyz <= center + half), dim=1)
    if kind == "sphere":
        return torch.linalg.norm(xyz - center, dim=1) <= sphere_radius
    height = cylinder_num[1] * particle_diameter
    d2 = (xyz[:, 0] - center[0]) ** 2 + (xyz[:, 1] - center[1]) ** 2
    return (d2 <= cylinder_radius ** 2) & (xyz[:, 2] >= center[2] - height / 2) & (xyz[:, 2] <= center[2] + height / 2)


def project_rigid(xyz, rigid_xyz, mask_inside, r, max_num_neighbors):
    """gm_fluid.py:1058-1105 / 1241-1289: the particles inside the body move onto their nearest rigid sample among the
    radius() edges (scatter_min over index-ordered edges).  Returns the new positions."""
    out = xyz.clone()
    if int(mask_inside.sum()) == 0:
        return out
    inside = xyz[mask_inside]
    e = radius(rigid_xyz, inside, r, max_num_neighbors=max_num_neighbors if max_num_neighbors else 10 ** 9)
    if e.size(1) == 0:
        return out
    row, col = e[0], e[1]
    dist2 = torch.sum((inside[row] - rigid_xyz[col]) ** 2, dim=1)
    _, argmin = scatter_min(dist2, row, dim=0, dim_size=inside.shape[0])
    has = argmin < row.numel()                    # (the reference assumes every inside particle has a neighbour)
    nearest = col[argmin[has]]
    moved = inside.clone()
    moved[has] = inside[has] + -(inside[has] - rigid_xyz[nearest])
    out[mask_inside] = moved
    return out
