"""Particle creation and emission between frames -- the host-side set-up around the solver tick.

Mirrors, with the reference's method and attribute names, what FluidNexus' entries call on `gm_dynamics.GaussianModel` before and
between the optimised / simulated frames (FD/entries_fluid_nexus/train_physical_particle.py:82,192,236,286,496;
future_simulation.py:102,132):

    create_particles_visual(model_args)                       gm_dynamics.py:510-555   random pillar of visual particles
    create_particles_hidden(model_args)                       gm_dynamics.py:558-609   lattice pillar of hidden particles + state
    prepare_emitter_points(model_args, is_future=False)       gm_dynamics.py:674-745   discs of emitter sites
    prepare_emitter_future_first_points(model_args)           gm_dynamics.py:747-788   stacked discs for the first future frames
    emit_new_particles(future_time_index=-1)                  gm_dynamics.py:844-976   append one tick's worth of new particles
    create_rigid_body()                                       gm_dynamics.py:612-672   surface samples of the cuboid / sphere / cylinder
    prepare_{hidden,visual,future_visual,rigid_body}_particles_for_rendering()   gm_dynamics.py:1636-1700   constant raw appearance
    get_* accessors, load_ply(path)                           gm_dynamics.py:200-338, 1702-1744   what the render pipes read

This is plain torch / numpy bookkeeping (no kernels): a few hundred points per frame.  What matters is that a run seeded like the
reference's produces the SAME particles: the sites are enumerated in the reference's order (x outermost, then y, then z), and the
global numpy / torch random streams are consumed by the same calls in the same order (np.random.uniform / random for the visual
pillar; torch.randperm for fractional emit ratios; torch.randperm + torch.rand_like for the "extra" visual particles).  Pinned
against the reference's own methods by tests/test_reference_emitter_golden.py (tools/make_emitter_golden.py).

`EmitterMixin` works on any object that carries the solver's state attributes (`_xyz, _estimate_xyz, _buoyancy, _force, _velocity,
_imass, _counts, _visual_xyz`) and a `dev` device; fluidnexus_b200.solver.PBFSolver inherits it.
"""
import numpy as np
import torch

# optim_args fields read by setup_constants (gm_dynamics.py:76-160) that steer the emission, with the defaults of
# FD/arguments/__init__.py:314-344
EMIT_DEFAULTS = dict(emit_ratio_hidden=1.32, emit_ratio_visual=1.32, extra_visual_ratio=0.0, extra_visual_num=0, extra_visual_y_min=0.16,
                     extra_visual_min_num=0, init_hidden_velocity=0.0)


def disc_sites(center_x, center_z, radius, delta, ys):
    """Sites (x, y, z) of a regular lattice of pitch `delta` inside the vertical cylinder of `radius` around (center_x, ., center_z),
    one layer per entry of `ys`, ordered x-major, then y, then z -- float64 [n, 3]."""
    xs = np.arange(center_x - radius, center_x + radius + delta, delta)   # (the stop is exclusive: one pitch is added)
    zs = np.arange(center_z - radius, center_z + radius + delta, delta)
    ys = np.asarray(ys, dtype=np.float64).reshape(-1)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    keep = (X - center_x) ** 2 + (Z - center_z) ** 2 <= radius ** 2
    return np.stack([X[keep], Y[keep], Z[keep]], axis=1).reshape(-1, 3)


def _f32(points, dev):
    return torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)).float().to(dev)


class EmitterMixin:
    """See the module docstring.  Needs: self.dev, self.scale_factor, self.alpha, gravity as self._gravity_t ([1,3] tensor) or
    self._gravity (3 floats), and the state attributes; emission settings are attributes with the reference's names
    (EMIT_DEFAULTS unless set)."""

    # -- settings -------------------------------------------------------------------------------------------------------------
    def setup_emitter(self, optim_args=None, **overrides):
        """The emission settings of setup_constants (gm_dynamics.py:76-160): attributes of `optim_args` where present, then
        keyword overrides, else the reference's defaults."""
        for k, v in EMIT_DEFAULTS.items():
            val = getattr(optim_args, k, v) if optim_args is not None else v
            setattr(self, k, overrides.get(k, val))
        self.emit_counter = 0

    def _emit_setting(self, name):
        return getattr(self, name, EMIT_DEFAULTS[name])

    def _gravity_row(self):
        g = getattr(self, "_gravity_t", None)
        if g is None:
            g = torch.tensor([float(c) for c in self._gravity], dtype=torch.float32, device=self.dev).reshape(1, 3)
        return g

    def _fresh_hidden_state(self, n):
        """State rows of n new hidden particles: estimate 0, buoyancy = gravity * alpha, no force, initial upward velocity, unit mass."""
        z3 = torch.zeros((n, 3), dtype=torch.float32, device=self.dev)
        vel = z3.clone()
        vel[:, 1] = self._emit_setting("init_hidden_velocity")
        buoy = torch.ones((n, 3), dtype=torch.float32, device=self.dev) * (self._gravity_row() * self.alpha)
        return dict(_estimate_xyz=z3, _buoyancy=buoy, _force=z3.clone(), _velocity=vel,
                    _imass=torch.ones((n, 1), dtype=torch.float32, device=self.dev))

    # -- first frame ----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def create_particles_visual(self, model_args):
        """A random pillar: `init_visual_num_pts` particles within `init_visual_radius_small_max` of the axis over the whole height
        plus `init_thick_visual_num_pts` within `init_visual_radius_max` over the upper part (render units, NOT scaled)."""
        n, nt = int(model_args.init_visual_num_pts), max(int(model_args.init_thick_visual_num_pts), 0)
        self.visual_x_mid, self.visual_z_mid = model_args.init_x_mid, model_args.init_z_mid
        y = np.random.uniform(model_args.init_visual_y_min, model_args.init_visual_y_max, (n, 1))
        if nt > 0:
            y = np.concatenate((y, np.random.uniform(model_args.init_visual_y_thick_min, model_args.init_visual_y_max, (nt, 1))), axis=0)
        r = np.random.random((n, 1)) * model_args.init_visual_radius_small_max
        if nt > 0:
            r = np.concatenate((r, np.random.random((nt, 1)) * model_args.init_visual_radius_max), axis=0)
        theta = np.random.random((n + nt, 1)) * 2 * np.pi
        pts = np.concatenate((r * np.cos(theta) + self.visual_x_mid, y, r * np.sin(theta) + self.visual_z_mid), axis=1)
        self._visual_xyz = _f32(pts, self.dev)
        self.visual_particles_created = True

    @torch.no_grad()
    def create_particles_hidden(self, model_args):
        """A lattice pillar of pitch `init_hidden_delta` (scaled units) with fresh state."""
        d = model_args.init_hidden_delta
        ys = np.arange(model_args.init_hidden_y_min, model_args.init_hidden_y_max, d)
        pts = disc_sites(model_args.init_x_mid, model_args.init_z_mid, model_args.init_hidden_radius_max, d, ys)
        self._xyz = _f32(pts * self.scale_factor, self.dev)
        n = self._xyz.shape[0]
        for name, t in self._fresh_hidden_state(n).items():
            setattr(self, name, t)
        self._counts = torch.zeros((n, 1), dtype=torch.float32, device=self.dev)
        self._particle_id = torch.arange(n, device=self.dev).unsqueeze(1)
        self._particle_id_max = n
        self.hidden_particles_created = True

    @torch.no_grad()
    def remove_invisible_bottom_visual_particles(self):
        """Drops the visual particles below y = -0.017 (render units), i.e. under the visible volume (gm_dynamics.py:1062-1070; called
        once before the first future frame).  Only the positions are filtered, like the reference."""
        keep = self._visual_xyz[:, 1] >= -0.017 * self.scale_factor
        if int(keep.sum()) < keep.shape[0]:
            self._visual_xyz = self._visual_xyz[keep]

    # -- rigid body -----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def create_rigid_body(self):
        """Surface samples of the rigid body (gm_dynamics.py:612-672) around `rigid_body_center` (scaled units): the shell of a
        `rigid_cuboid_num` lattice of pitch `rigid_particle_diameter`, `rigid_sphere_num` uniformly random points on a sphere, or
        `rigid_cylinder_num` = (around, along) points on a cylinder about the z axis.  Reads the attributes of the reference's
        setup_constants (`rigid_body`, `rigid_particle_diameter`, ...); writes `_rigid_xyz` [n,3] and `_rigid_imass` [n,1] = 0."""
        diam = self.rigid_particle_diameter
        if self.rigid_body == "cuboid":
            nx, ny, nz = self.rigid_cuboid_num
            axes = [np.arange(n) * diam - n // 2 * diam for n in (nx, ny, nz)]
            I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
            shell = (I == 0) | (I == nx - 1) | (J == 0) | (J == ny - 1) | (K == 0) | (K == nz - 1)   # faces only, no interior
            pts = np.stack([axes[0][I[shell]], axes[1][J[shell]], axes[2][K[shell]]], axis=1)
        elif self.rigid_body == "sphere":
            n, radius = self.rigid_sphere_num, self.rigid_sphere_radius
            phi = np.random.uniform(0, 2 * np.pi, n)
            theta = np.arccos(np.random.uniform(-1, 1, n))           # uniform in cos(theta): uniform on the sphere
            pts = np.vstack((radius * np.sin(theta) * np.cos(phi), radius * np.sin(theta) * np.sin(phi), radius * np.cos(theta))).T
        elif self.rigid_body == "cylinder":
            around, along = self.rigid_cylinder_num
            theta = np.repeat(np.arange(around) * 2 * np.pi / around, along)
            z = np.tile((np.arange(along) - along / 2) * diam, around)
            pts = np.stack([self.rigid_cylinder_radius * np.cos(theta), self.rigid_cylinder_radius * np.sin(theta), z], axis=1)
        else:
            raise ValueError(f"rigid_body must be cuboid, sphere or cylinder (got {self.rigid_body!r})")
        center = torch.as_tensor(self.rigid_body_center, dtype=torch.float32).to(self.dev)
        self._rigid_xyz = _f32(pts, self.dev) + center
        self._rigid_imass = torch.zeros((pts.shape[0], 1), dtype=torch.float32, device=self.dev)

    # -- what the render pipes read (renderer.py / FD/renderer/pipe_*.py): the reference's accessors, gm_dynamics.py:200-338 --------
    # raw tensors -> activated attributes: exp (scales), sigmoid (opacity), normalise (rotation); colours and positions as stored
    active_sh_degree = 0
    scaling_activation = staticmethod(torch.exp)
    opacity_activation = staticmethod(torch.sigmoid)
    rotation_activation = staticmethod(torch.nn.functional.normalize)
    get_xyz = property(lambda s: s._xyz)
    get_estimate_xyz = property(lambda s: s._estimate_xyz)
    get_force = property(lambda s: s._force)
    get_velocity = property(lambda s: s._velocity)
    get_imass = property(lambda s: s._imass)
    get_color_dummy = property(lambda s: s._color_dummy)
    get_scaling_dummy = property(lambda s: s.scaling_activation(s._scales_dummy))
    get_rotation_dummy = property(lambda s: s.rotation_activation(s._rotation_dummy))
    get_opacity_dummy = property(lambda s: s.opacity_activation(s._opacity_dummy))
    get_visual_xyz = property(lambda s: s._visual_xyz)
    get_visual_color = property(lambda s: s._visual_color)
    get_visual_scaling = property(lambda s: s.scaling_activation(s._visual_scales))
    get_visual_rotation = property(lambda s: s.rotation_activation(s._visual_rotation))
    get_visual_opacity = property(lambda s: s.opacity_activation(s._visual_opacity))
    get_rigid_xyz = property(lambda s: s._rigid_xyz)
    get_rigid_color = property(lambda s: s._rigid_color)
    get_rigid_scaling = property(lambda s: s.scaling_activation(s._rigid_scales))
    get_rigid_rotation = property(lambda s: s.rotation_activation(s._rigid_rotation))
    get_rigid_opacity = property(lambda s: s.opacity_activation(s._rigid_opacity))
    get_gs_xyz = property(lambda s: s._gs_xyz)
    get_gs_color = property(lambda s: s._gs_color)
    get_gs_scaling = property(lambda s: s.scaling_activation(s._gs_scales))
    get_gs_rotation = property(lambda s: s.rotation_activation(s._gs_rotation))
    get_gs_opacity = property(lambda s: s.opacity_activation(s._gs_opacity))

    def load_ply(self, path):
        """The frozen background set that render_dynamics concatenates behind the particles (gm_dynamics.py:1702-1744): the point
        cloud the background stage wrote, raw attributes, x / y un-negated (fluidnexus_b200/io.py:load_background_ply)."""
        from . import io as IO
        d = IO.load_background_ply(path)
        for mine, key in (("_gs_xyz", "xyz"), ("_gs_color", "color"), ("_gs_opacity", "opacity"), ("_gs_scales", "scaling"), ("_gs_rotation", "rotation")):
            setattr(self, mine, torch.from_numpy(np.ascontiguousarray(d[key], dtype=np.float32)).to(self.dev))
        self.active_sh_degree = 0
        return int(self._gs_xyz.shape[0])

    # -- raw attributes the render pipes read for particle sets that have no trained appearance ------------------------------------
    # (gm_dynamics.py:1636-1700; constants of setup_constants :158-160: colour 0.7, log-scale -5.9, opacity 0.1; rigid body: 0.9 / -5.5 / 0.3)
    constant_color, constant_scale, constant_opacity = 0.7, -5.9, 0.1

    def _constant_appearance(self, n, color, log_scale, opacity):
        """Raw (pre-activation) colour [n,1], log-scales [n,3], identity quaternions [n,4], opacity logits [n,1]."""
        full = lambda cols, v: torch.zeros((n, cols), dtype=torch.float32, device=self.dev) + v
        rot = full(4, 0.0)
        rot[:, 0] = 1.0
        op = opacity * torch.ones((n, 1), dtype=torch.float32, device=self.dev)
        return full(1, color), full(3, log_scale), rot, torch.log(op / (1 - op))          # inv_sigmoid, general_utils.py:10-11

    def prepare_hidden_particles_for_rendering(self):
        assert self._xyz.shape[0] > 0, "No hidden particles to render"
        self._color_dummy, self._scales_dummy, self._rotation_dummy, self._opacity_dummy = self._constant_appearance(
            self._xyz.shape[0], self.constant_color, self.constant_scale, self.constant_opacity)

    def prepare_visual_particles_for_rendering(self):
        assert self._visual_xyz.shape[0] > 0, "No visual particles to render"
        self._visual_color, self._visual_scales, self._visual_rotation, self._visual_opacity = self._constant_appearance(
            self._visual_xyz.shape[0], self.constant_color, self.constant_scale, self.constant_opacity)

    def prepare_future_visual_particles_for_rendering(self, use_level_two_future=False):
        """Future simulation: particles emitted since the last call get the constant appearance, the earlier ones keep theirs (the
        level-two result) when use_level_two_future, else all are reset."""
        if not use_level_two_future:
            return self.prepare_visual_particles_for_rendering()
        new = self._constant_appearance(self._visual_xyz.shape[0] - self._visual_color.shape[0], self.constant_color, self.constant_scale,
                                        self.constant_opacity)
        for name, t in zip(("_visual_color", "_visual_scales", "_visual_rotation", "_visual_opacity"), new):
            setattr(self, name, torch.cat((getattr(self, name), t), dim=0))

    def prepare_rigid_body_particles_for_rendering(self):
        assert self._rigid_xyz.shape[0] > 0, "No rigid body particles to render"
        self._rigid_color, self._rigid_scales, self._rigid_rotation, self._rigid_opacity = self._constant_appearance(
            self._rigid_xyz.shape[0], 0.9, -5.5, 0.3)

    # -- emitter sites --------------------------------------------------------------------------------------------------------
    def _emitter_geometry(self, model_args):
        dh, dv = model_args.emitter_hidden_delta, model_args.emitter_visual_delta
        return (dh, dv, model_args.init_x_mid, model_args.init_z_mid, dv * model_args.emitter_visual_radius_ratio,
                dh * model_args.emitter_hidden_radius_ratio)

    @torch.no_grad()
    def prepare_emitter_points(self, model_args, is_future=False):
        """One disc of sites each for the visual and the hidden particles at the emitter heights (render units); for future
        simulation the visual disc sits half a radius lower."""
        dh, dv, cx, cz, rv, rh = self._emitter_geometry(model_args)
        self.hidden_delta_offset, self.visual_delta_offset = dh, dv
        y_vis = model_args.emitter_center_y_visual - rv / 2 if is_future else model_args.emitter_center_y_visual
        self.visual_emitter_points = _f32(disc_sites(cx, cz, rv, dv, [y_vis]), self.dev)
        self.hidden_emitter_points = _f32(disc_sites(cx, cz, rh, dh, [model_args.emitter_center_y_hidden]), self.dev)

    @torch.no_grad()
    def prepare_emitter_future_first_points(self, model_args):
        """Stacks of discs (one diameter high) that the first two future frames emit at once."""
        dh, dv, cx, cz, rv, rh = self._emitter_geometry(model_args)
        yv0, yh0 = model_args.emitter_center_y_visual, model_args.emitter_center_y_hidden
        self.visual_emitter_first_points = _f32(disc_sites(cx, cz, rv, dv, np.arange(yv0, yv0 + rv * 2 + dv, dv)), self.dev)
        self.hidden_emitter_first_points = _f32(disc_sites(cx, cz, rh, dh, np.arange(yh0, yh0 + rh * 2 + dh, dh)), self.dev)

    # -- per tick -------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _copies_of(sites, ratio, scale):
        """`ratio` copies of the site set: int(ratio) whole copies, then a random subset holding the fractional part of one more."""
        whole, frac = int(ratio), ratio - int(ratio)
        out = [sites.clone() * scale for _ in range(whole)]
        if frac > 0:
            cand = sites.clone() * scale
            out.append(cand[torch.randperm(cand.shape[0])[: int(frac * cand.shape[0])]])
        return out

    def _extra_visual(self, count_of):
        """Copies of randomly chosen visual particles above `extra_visual_y_min`, jittered by 5 % of the visual pitch."""
        sf = self.scale_factor
        high = self._visual_xyz[self._visual_xyz[:, 1] > self._emit_setting("extra_visual_y_min") * sf]
        pick = torch.randperm(high.shape[0])[: count_of(high.shape[0])]
        chosen = high[pick] / sf
        jitter = self.visual_delta_offset * (torch.rand_like(chosen) - 0.5) * 0.05
        return (chosen + jitter) * sf

    @torch.no_grad()
    def emit_new_particles(self, future_time_index=-1):
        self.emit_counter = getattr(self, "emit_counter", 0) + 1
        sf = self.scale_factor
        if 0 <= future_time_index < 2:
            new_hidden = [self.hidden_emitter_first_points.clone() * sf]
            new_visual = [self.visual_emitter_first_points.clone() * sf]
        else:
            new_hidden = self._copies_of(self.hidden_emitter_points, self._emit_setting("emit_ratio_hidden"), sf)
            new_visual = self._copies_of(self.visual_emitter_points, self._emit_setting("emit_ratio_visual"), sf)
            ratio, min_num = self._emit_setting("extra_visual_ratio"), self._emit_setting("extra_visual_min_num")
            if ratio > 0.0:
                new_visual.append(self._extra_visual(lambda n_high: max(int(n_high * ratio), min_num)))
            if self._emit_setting("extra_visual_num") > 0:
                new_visual.append(self._extra_visual(lambda n_high: self._emit_setting("extra_visual_num")))
        if new_hidden:
            add = torch.cat(new_hidden, dim=0)
            n_new = add.shape[0]
            self._xyz = torch.cat((self._xyz, add), dim=0)
            for name, t in self._fresh_hidden_state(n_new).items():
                setattr(self, name, torch.cat((getattr(self, name), t), dim=0))
            # the solver counts of ALL particles start over (gm_dynamics.py:962)
            self._counts = torch.zeros((self._xyz.shape[0], 1), dtype=torch.float32, device=self.dev)
            if hasattr(self, "_particle_id"):
                ids = torch.arange(self._particle_id_max, self._particle_id_max + n_new, device=self.dev).unsqueeze(1)
                self._particle_id = torch.cat((self._particle_id, ids), dim=0)
                self._particle_id_max += n_new
        if new_visual:
            add = torch.cat(new_visual, dim=0)
            self._visual_xyz = add if self._visual_xyz.shape[0] == 0 else torch.cat((self._visual_xyz, add), dim=0)
