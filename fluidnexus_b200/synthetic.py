"""Seeded synthetic scenes in the reference's conventions.

Cameras reproduce FD/scene/camera.py:90-110 + FD/utils/graphics_utils.py:24-60 exactly: `world_view_transform`
and `full_proj_transform` are the row-major storage of the TRANSPOSED 4x4 matrices, z_near 0.01, z_far 100.
Particle/Gaussian statistics follow FD/gaussian_splatting/gm_fluid.py (init values :1531-1535, lattice :489-529,
scale_factor 100 :122) as collected in SURVEY.md 8(d).  Everything is generated with numpy on the host from a
seed, so CPU tests, the oracle, and the GPU benchmark see bit-identical inputs.
"""
import math
from dataclasses import dataclass, field

import numpy as np
import torch

PLUME_CENTER = np.array([0.34, 0.3, -0.225], dtype=np.float64)  # emitter centre, gm_fluid.py:597
SCALE_FACTOR = 100.0  # gm_fluid.py:122
SECS = 0.033


def world_to_view(R, t, translate=np.zeros(3), scale=1.0):
    """graphics_utils.py:24-36 (get_world_2_view2)."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    C2W[:3, 3] = (C2W[:3, 3] + translate) * scale
    return np.float32(np.linalg.inv(C2W))


def projection_matrix(z_near, z_far, fovX, fovY):
    """graphics_utils.py:39-60 (get_projection_matrix); note P[2,2] uses (far+near)/(far-near)."""
    tx, ty = math.tan(fovX / 2), math.tan(fovY / 2)
    top, right = ty * z_near, tx * z_near
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * z_near / (right - left)
    P[1, 1] = 2.0 * z_near / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = (z_far + z_near) / (z_far - z_near)
    P[2, 3] = -(z_far * z_near) / (z_far - z_near)
    return P


class SyntheticCamera:
    """The attributes of FD/scene/camera.py:Camera that the render pipes and entries read."""

    def __init__(self, R, T, FoVx, FoVy, width, height, name, timestamp=0.0, device="cpu"):
        self.R, self.T, self.FoVx, self.FoVy = R, T, FoVx, FoVy
        self.image_width, self.image_height = width, height
        self.image_name, self.timestamp = name, timestamp
        self.z_far, self.z_near = 100.0, 0.01
        self.world_view_transform = torch.tensor(world_to_view(R, T)).transpose(0, 1).to(device)
        self.projection_matrix = projection_matrix(self.z_near, self.z_far, FoVx, FoVy).transpose(0, 1).to(device)
        self.full_proj_transform = (
            self.world_view_transform.unsqueeze(0).bmm(self.projection_matrix.unsqueeze(0))).squeeze(0)
        self.camera_center = self.world_view_transform.inverse()[3, :3]
        self.original_image = None       # CPU tensor [C,H,W], like the reference (uploaded every iteration)
        self.original_image_real = None

    def to(self, device):
        for k in ("world_view_transform", "projection_matrix", "full_proj_transform", "camera_center"):
            setattr(self, k, getattr(self, k).to(device))
        return self


def look_at_camera(eye, target, fov_x, width, height, name, device="cpu"):
    """COLMAP-style camera (x right, y down, z forward); R is camera-to-world as the reference stores it."""
    z = target - eye
    z = z / np.linalg.norm(z)
    down = np.array([0.0, -1.0, 0.0])
    x = np.cross(down, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=1)  # columns = camera axes in world space
    T = -R.transpose() @ eye
    fov_y = 2 * math.atan(math.tan(fov_x / 2) * height / width)
    return SyntheticCamera(R, T, fov_x, fov_y, width, height, name, device=device)


def make_cameras(n_views=5, size=512, radius=1.0, arc_deg=120.0, fov_x=0.69, device="cpu", height=None):
    """`n_views` cameras on a horizontal arc around the plume centre (README.md:48 of the reference)."""
    cams = []
    h = size if height is None else height
    for k in range(n_views):
        ang = math.radians(-arc_deg / 2 + arc_deg * (k / max(1, n_views - 1))) if n_views > 1 else 0.0
        eye = PLUME_CENTER + radius * np.array([math.sin(ang), 0.0, math.cos(ang)])
        cams.append(look_at_camera(eye, PLUME_CENTER, fov_x, size, h, f"view_{k:02d}", device=device))
    return cams


@dataclass
class GaussianSet:
    """Activated Gaussian attributes as the render pipes hand them to the rasterizer."""
    xyz: np.ndarray        # [P,3] render units
    scales: np.ndarray     # [P,3] (already exp-ed)
    rotations: np.ndarray  # [P,4] normalised, (r,x,y,z)
    opacity: np.ndarray    # [P,1] in (0,1)
    colors: np.ndarray     # [P,C]

    @property
    def P(self):
        return self.xyz.shape[0]

    def torch(self, device):
        return {k: torch.from_numpy(np.ascontiguousarray(getattr(self, k), dtype=np.float32)).to(device)
                for k in ("xyz", "scales", "rotations", "opacity", "colors")}


def cat_sets(a, b):
    return GaussianSet(*[np.concatenate([getattr(a, k), getattr(b, k)], 0)
                         for k in ("xyz", "scales", "rotations", "opacity", "colors")])


def _rotations(rng, n):
    q = np.zeros((n, 4))
    q[:, 0] = 1.0
    q += rng.normal(0, 0.1, (n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def fluid_gaussians(P, channels, seed=0, radius=0.1, height=0.6, log_scale=-5.9):
    """Fluid particles: uniform in a cylinder r <= 0.1, y in [0,0.6] around the plume axis."""
    rng = np.random.default_rng(seed)
    r = radius * np.sqrt(rng.uniform(0, 1, P))
    th = rng.uniform(0, 2 * math.pi, P)
    xyz = np.stack([PLUME_CENTER[0] + r * np.cos(th), rng.uniform(0, height, P), PLUME_CENTER[2] + r * np.sin(th)], 1)
    scales = np.exp(log_scale + rng.uniform(-0.3, 0.3, (P, 3)))
    opacity = rng.uniform(0.05, 0.3, (P, 1))
    grey = rng.uniform(0.4, 0.9, (P, 1))
    colors = np.repeat(grey, channels, 1)  # grey particles, repeated to RGB like pipe_dynamics.py:118-120
    return GaussianSet(xyz, scales, _rotations(rng, P), opacity, colors)


def _in_frustum_count(xyz, cams, margin=1.0):
    """number of cameras whose image (scaled by `margin`) contains each point, z > 0.2"""
    cnt = np.zeros(xyz.shape[0], dtype=np.int32)
    for cam in cams:
        Vm = cam.world_view_transform.cpu().numpy().astype(np.float64)  # transposed storage: p_view = p_h @ Vm
        ph = np.concatenate([xyz, np.ones((xyz.shape[0], 1))], 1) @ Vm
        z = ph[:, 2]
        ok = z > 0.2
        tx, ty = math.tan(cam.FoVx / 2) * margin, math.tan(cam.FoVy / 2) * margin
        with np.errstate(divide="ignore", invalid="ignore"):
            ok &= (np.abs(ph[:, 0] / z) < tx) & (np.abs(ph[:, 1] / z) < ty)
        cnt += ok
    return cnt


def background_gaussians(P, channels, seed=1, cams=None, min_views=3):
    """Frozen background: uniform on a box shell 0.5-1.5 from the plume centre, kept only where at least
    `min_views` of the cameras see it (a reconstructed static scene lives in the cameras' common frustum)."""
    rng = np.random.default_rng(seed)
    cams = make_cameras(5, 512) if cams is None else cams
    chunks, have = [], 0
    while have < P:
        n = max(4 * (P - have), 1024)
        d = rng.normal(size=(n, 3))
        d /= np.abs(d).max(axis=1, keepdims=True)  # on the unit cube surface
        cand = PLUME_CENTER + d * rng.uniform(0.5, 1.5, (n, 1))
        cand = cand[_in_frustum_count(cand, cams) >= min(min_views, len(cams))]
        chunks.append(cand)
        have += cand.shape[0]
    xyz = np.concatenate(chunks, 0)[:P]
    scales = np.exp(rng.uniform(-5.0, -3.0, (P, 3)))
    opacity = rng.uniform(0.2, 0.9, (P, 1))
    colors = rng.uniform(0, 1, (P, channels))
    return GaussianSet(xyz, scales, _rotations(rng, P), opacity, colors)


BALL_CENTER = PLUME_CENTER + np.array([0.0, 0.32, 0.0])   # the rigid ball hangs in the plume (FluidNexus-Ball)
BALL_RADIUS = 0.05


def ball_gaussians(P, channels, seed=4, center=BALL_CENTER, radius=BALL_RADIUS):
    """The ball of the FluidNexus-Ball scenes as the frozen background set holds it after train_background: an opaque shell
    of small Gaussians (gm_dynamics.load_ply, FD/gaussian_splatting/gm_dynamics.py:1702-1744; there is no separate rigid
    object at training time)."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(P, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    xyz = center + d * (radius * (1.0 - 0.08 * rng.uniform(0, 1, (P, 1))))
    scales = np.exp(rng.uniform(-5.6, -4.6, (P, 3)))
    opacity = rng.uniform(0.6, 0.95, (P, 1))
    base = np.array([0.8, 0.25, 0.2])[:channels] if channels == 3 else np.array([0.6])
    colors = np.clip(base + rng.normal(0, 0.05, (P, channels)), 0, 1)
    return GaussianSet(xyz, scales, _rotations(rng, P), opacity, colors)


def random_gaussians(P, channels, seed=0, spread=0.25, log_scale=(-5.0, -3.2)):
    """Small generic test scene in front of the default cameras."""
    rng = np.random.default_rng(seed)
    xyz = PLUME_CENTER + rng.uniform(-spread, spread, (P, 3))
    scales = np.exp(rng.uniform(log_scale[0], log_scale[1], (P, 3)))
    opacity = rng.uniform(0.05, 0.95, (P, 1))
    colors = rng.uniform(0, 1, (P, channels))
    return GaussianSet(xyz, scales, _rotations(rng, P), opacity, colors)


@dataclass
class HiddenParticles:
    """Hidden (simulation) particles in SCALED units (x100), gm_fluid.py state names in brackets."""
    xyz: np.ndarray           # [N,3] (_xyz)
    velocity: np.ndarray      # [N,3]
    estimate_xyz: np.ndarray  # [N,3] (_estimate_xyz = _xyz + secs*v)
    buoyancy: np.ndarray      # [N,3] (_buoyancy = gravity*alpha)
    force: np.ndarray         # [N,3] (_force)
    imass: np.ndarray         # [N,1] (_imass)
    secs: float = SECS
    extra: dict = field(default_factory=dict)

    @property
    def N(self):
        return self.xyz.shape[0]


def hidden_lattice(N, seed=2, spacing=0.9, jitter=0.05, secs=SECS, buoyancy=(0.0, 0.0, 0.0)):
    """N lattice particles (spacing 0.9 scaled units, gm_fluid.py:493) filling a column around the plume axis,
    jittered by U(+-jitter); velocities N(0,5) + 30 y."""
    rng = np.random.default_rng(seed)
    c = PLUME_CENTER * SCALE_FACTOR
    nx = nz = int(math.ceil(22.0 / spacing))
    ny = int(math.ceil(62.0 / spacing))
    while nx * ny * nz < 2 * N:
        nx += 1; nz += 1; ny += 2
    gx = (np.arange(nx) - (nx - 1) / 2) * spacing + c[0]
    gy = np.arange(ny) * spacing
    gz = (np.arange(nz) - (nz - 1) / 2) * spacing + c[2]
    X, Y, Z = np.meshgrid(gx, gy, gz, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    # keep the N lattice sites closest to the axis-aligned column (deterministic)
    score = np.hypot(pts[:, 0] - c[0], pts[:, 2] - c[2]) + 1e-3 * np.maximum(pts[:, 1] - 60.0, 0) * 1e3
    keep = np.argsort(score, kind="stable")[:N]
    keep.sort()
    xyz = pts[keep] + rng.uniform(-jitter, jitter, (N, 3))
    vel = rng.normal(0, 5.0, (N, 3)) + np.array([0.0, 30.0, 0.0])
    return HiddenParticles(xyz=xyz, velocity=vel, estimate_xyz=xyz + secs * vel,
                           buoyancy=np.tile(np.asarray(buoyancy, dtype=np.float64), (N, 1)), force=np.zeros((N, 3)),
                           imass=np.ones((N, 1)), secs=secs)


def cube_lattice(n_side=32, seed=3, spacing=0.9, jitter=0.05, secs=SECS):
    """BASELINE config 1: a 32^3 = 32768-particle lattice."""
    rng = np.random.default_rng(seed)
    g = (np.arange(n_side) - (n_side - 1) / 2) * spacing
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    xyz = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1) + PLUME_CENTER * SCALE_FACTOR
    N = xyz.shape[0]
    xyz = xyz + rng.uniform(-jitter, jitter, (N, 3))
    vel = rng.normal(0, 5.0, (N, 3)) + np.array([0.0, 30.0, 0.0])
    return HiddenParticles(xyz=xyz, velocity=vel, estimate_xyz=xyz + secs * vel, buoyancy=np.zeros((N, 3)),
                           force=np.zeros((N, 3)), imass=np.ones((N, 1)), secs=secs)


def raster_inputs(gs: GaussianSet, cam: SyntheticCamera, bg=None):
    """Keyword arguments in the order of `_C.rasterize_gaussians` (R3/rasterize_points.h:18-37) as numpy."""
    C = gs.colors.shape[1]
    return dict(
        bg=np.zeros(C, np.float32) if bg is None else np.asarray(bg, np.float32),
        means3D=gs.xyz.astype(np.float32), colors=gs.colors.astype(np.float32), opacities=gs.opacity.astype(np.float32),
        scales=gs.scales.astype(np.float32), rotations=gs.rotations.astype(np.float32), scale_modifier=1.0,
        view=cam.world_view_transform.cpu().numpy().astype(np.float32),
        proj=cam.full_proj_transform.cpu().numpy().astype(np.float32),
        tan_fov_x=math.tan(cam.FoVx * 0.5), tan_fov_y=math.tan(cam.FoVy * 0.5), H=cam.image_height, W=cam.image_width,
    )
