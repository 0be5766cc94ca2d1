"""No-grad PBF solver tick on libfnx.

Host-side mirror of the simulation half of FD/gaussian_splatting/gm_fluid.py's GaussianModel -- the part that runs
between the optimised frames (FD/entries_fluid_nexus/train_physical_particle.py:206-216,288) and in all of
future_simulation.py:135-162:

    guess_hidden_particles(stable, use_wind)        gm_fluid.py:809-844
    update_solver_counts()                          gm_fluid.py:893-894
    project_gas_constraints()                       gm_fluid.py:896-1021   (solver_iterations times per tick)
    confirm_guess_hidden_particles()                gm_fluid.py:1160-1175
    update_visual_particles()                       gm_fluid.py:1197-1239
    remove_invalid_particles()                      gm_fluid.py:864-891
    create_particles_* / prepare_emitter_points / emit_new_particles   gm_dynamics.py:510-609, 674-788, 844-976 (emitter.py: host-side
                                                    set-up between frames, same particles as the reference under the same seeds)

Same method names, same in-place state updates (attributes `_xyz, _estimate_xyz, _velocity, _force, _buoyancy, _imass,
_counts, _visual_xyz`), no CPU fallback.  One solver iteration is 9 kernel launches (grid build, neighbour count, two
gather passes) instead of the reference's ~40 torch kernels over a materialised edge list plus 20 `.item()` syncs for its
log dictionary; `project_gas_constraints(stats=True)` returns that dictionary's per-node entries on request only.
"""
import ctypes as C

import torch

from . import _lib as L
from .emitter import EmitterMixin


class PBFSolver(EmitterMixin):
    """State + the reference's solver methods.  Constants default to gm_fluid.py:98-106 / FD/arguments/__init__.py."""

    def __init__(self, xyz, velocity=None, imass=None, visual_xyz=None, H=2.0, p0=1.5, k=10.0, KNN_K=100, secs=0.033, alpha=-0.2,
                 buoyancy_max_y=0.0, buoyancy_decay_rate=0.0, scale_factor=100.0, gravity=(0.0, -9.8, 0.0), wind_force=(0.0, 0.0, 0.0),
                 wind_power=1.0, min_neighbors=-1, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("PBFSolver runs on CUDA tensors only (no CPU fallback)")
        f = lambda t: torch.as_tensor(t, dtype=torch.float32).to(dev).contiguous().clone()
        self.dev = dev
        self._xyz = f(xyz) if xyz is not None else torch.zeros((0, 3), device=dev)   # None: start empty (create_particles_hidden / emitter)
        N = self._xyz.size(0)
        self._velocity = f(velocity) if velocity is not None else torch.zeros((N, 3), device=dev)
        self._imass = f(imass).reshape(N, 1) if imass is not None else torch.ones((N, 1), device=dev)
        self._estimate_xyz = self._xyz.clone()
        self._force = torch.zeros((N, 3), device=dev)
        self._buoyancy = torch.zeros((N, 3), device=dev)
        self._counts = torch.zeros((N, 1), device=dev)
        self._visual_xyz = f(visual_xyz) if visual_xyz is not None else torch.zeros((0, 3), device=dev)
        self.H, self.p0, self.k, self.KNN_K, self._secs, self.alpha = float(H), float(p0), float(k), int(KNN_K), float(secs), float(alpha)
        self.buoyancy_max_y, self.buoyancy_decay_rate, self.scale_factor = float(buoyancy_max_y), float(buoyancy_decay_rate), float(scale_factor)
        self.RELAXATION, self.K_P, self.E_P, self.DQ_P = 0.01, 0.2, 4, 0.25
        self.min_neighbors = int(min_neighbors)
        self._gravity = (C.c_float * 3)(*gravity)
        self._wind = (C.c_float * 3)(*wind_force)
        self.wind_force_max, self.wind_power = float(max(wind_force)), float(wind_power)
        self._scratch_n = -1

    # -- scratch (re-made when the particle count changes) -----------------------------------------------------
    def _scratch(self):
        N, V = self._xyz.size(0), self._visual_xyz.size(0)
        if self._scratch_n != (N, V):
            lib = L.lib()
            self._grid = torch.empty(lib.fnx_grid_bytes(max(N, 1)), dtype=torch.uint8, device=self.dev)
            self._kth = torch.empty(max(N, 1), dtype=torch.int32, device=self.dev)
            self._kthV = torch.empty(max(V, 1), dtype=torch.int32, device=self.dev)
            self._lambda = torch.empty(max(N, 1), device=self.dev)
            self._nlen = torch.empty(max(N, 1), device=self.dev)
            self._pratio = torch.empty(max(N, 1), device=self.dev)
            self._scratch_n = (N, V)

    def _st(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    # -- the reference's methods --------------------------------------------------------------------------------
    @torch.no_grad()
    def guess_hidden_particles(self, stable=False, use_wind=False):
        cur_secs, cur_alpha = (0.01, -1.0) if stable else (self._secs, self.alpha)
        N = self._xyz.size(0)
        with torch.cuda.device(self.dev):
            L.check(L.lib().fnx_pbf_guess_hidden(N, self._xyz.data_ptr(), self._velocity.data_ptr(), self._buoyancy.data_ptr(),
                                                 self._force.data_ptr(), self._estimate_xyz.data_ptr(), self._counts.data_ptr(), self._gravity,
                                                 cur_alpha, cur_secs, self.buoyancy_max_y, self.scale_factor, self.buoyancy_decay_rate,
                                                 int(bool(use_wind)), self._wind, self.wind_power, self.wind_force_max, self._st()))

    def update_solver_counts(self):
        self._counts += 1.0

    @torch.no_grad()
    def project_gas_constraints(self, stats=False):
        self._scratch()
        N = self._estimate_xyz.size(0)
        with torch.cuda.device(self.dev):
            before = self._estimate_xyz.clone() if stats else None
            L.check(L.lib().fnx_pbf_project_gas_constraints(
                self._grid.data_ptr(), self._estimate_xyz.data_ptr(), N, self._imass.data_ptr(), self._velocity.data_ptr(),
                self._force.data_ptr(), self._counts.data_ptr(), self.H, self.p0, self.k, self.KNN_K, self.RELAXATION, self.K_P, self.E_P,
                self.DQ_P, self._kth.data_ptr(), self._lambda.data_ptr(), self._nlen.data_ptr(), self._pratio.data_ptr(), self._st()))
        if not stats:
            return None
        # the per-node entries of the reference's log dictionary (gm_fluid.py:998-1019); one sync, on request only
        m = lambda t: float(t[:N].mean())
        return {"velocity": m(self._velocity), "xyz": m(self._xyz), "estimate_xyz": m(self._estimate_xyz), "p_ratio": m(self._pratio),
                "pi": m(self._pratio) * self.p0, "lambdas": m(self._lambda), "estimate_xyz_delta": m(self._estimate_xyz - before),
                "elapsed_time": 0.0}

    @torch.no_grad()
    def confirm_guess_hidden_particles(self):
        with torch.cuda.device(self.dev):
            L.check(L.lib().fnx_pbf_confirm_guess(self._xyz.size(0), self._xyz.data_ptr(), self._estimate_xyz.data_ptr(),
                                                  self._velocity.data_ptr(), self._secs, self._st()))

    # the reference has a second method with the same body (gm_fluid.py:1177-1191), called after a frame's optimisation
    confirm_guess_hidden_particles_wo_velocity = confirm_guess_hidden_particles

    @torch.no_grad()
    def confirm_guess_hidden_particles_from_nn(self, estimate_xyz_nn):
        """After a frame's optimisation (train_physical_particle.py:432): the optimised positions (render units, the trainable tensor
        of the fused step: FrameState.e) become the solver's estimate (gm_fluid.py:1193-1195)."""
        self._estimate_xyz = (estimate_xyz_nn.detach().to(self.dev, torch.float32) * self.scale_factor).contiguous()

    @torch.no_grad()
    def update_visual_xyz_from_nn(self):
        """The visual particles move with the optimised hidden velocities (P1 forward once more, gm_fluid.py:1339-1340); call after
        confirm_guess_hidden_particles_from_nn."""
        from . import physics
        if self._visual_xyz.size(0):
            self._visual_xyz = physics.visual_advect(self._estimate_xyz, self._xyz, self._visual_xyz, self.H, self._secs, self.KNN_K).detach()

    # what fluidnexus_b200.step.FrameState reads from its `hidden` argument: FrameState(sol, sol._visual_xyz, fluid, background)
    # builds the optimisation state of the current frame straight from the solver (training_setup_current, gm_fluid.py:330-334)
    N = property(lambda self: int(self._xyz.size(0)))
    xyz = property(lambda self: self._xyz)
    estimate_xyz = property(lambda self: self._estimate_xyz)
    buoyancy = property(lambda self: self._buoyancy)
    force = property(lambda self: self._force)
    imass = property(lambda self: self._imass)

    @torch.no_grad()
    def update_visual_particles(self):
        V, N = self._visual_xyz.size(0), self._estimate_xyz.size(0)
        if V == 0:
            return
        self._scratch()
        with torch.cuda.device(self.dev):
            L.check(L.lib().fnx_pbf_update_visual(self._grid.data_ptr(), self._estimate_xyz.data_ptr(), self._velocity.data_ptr(), N,
                                                  self._visual_xyz.data_ptr(), V, self.KNN_K, self.H, self._secs, self._kthV.data_ptr(),
                                                  self._st()))

    @torch.no_grad()
    def remove_invalid_particles(self):
        """Drops particles with fewer than `min_neighbors` neighbours in radius_graph(xyz, H, loop=False) (torch_cluster's
        default cap of 32 neighbours applies, as in the reference call)."""
        if self.min_neighbors < 0:
            return
        self._scratch()
        N = self._xyz.size(0)
        degree = torch.empty(max(N, 1), dtype=torch.int32, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(L.lib().fnx_radius_graph_degree(self._grid.data_ptr(), self._xyz.data_ptr(), N, self.H, 0, 32, self._kth.data_ptr(),
                                                    degree.data_ptr(), self._st()))
        mask = degree[:N] >= self.min_neighbors
        if not bool(mask.all()):
            for name in ("_xyz", "_estimate_xyz", "_buoyancy", "_force", "_velocity", "_imass", "_counts", "_particle_id"):
                if hasattr(self, name):      # (_particle_id exists once particles were created / emitted / loaded here)
                    setattr(self, name, getattr(self, name)[mask].contiguous())

    # -- rigid coupling (gm_fluid.py:1023-1105, 1241-1289) ---------------------------------------------------------
    def set_rigid_body(self, kind, center, rigid_xyz, cuboid_num=None, particle_radius=None, sphere_radius=None, cylinder_radius=None,
                       cylinder_num=None):
        """kind in {"cuboid", "sphere", "cylinder"}; center and rigid_xyz in scaled units.  cuboid: edge = num * 2 *
        particle_radius per axis; cylinder: height = cylinder_num[1] * 2 * particle_radius (check_inside_rigid_body)."""
        assert kind in ("cuboid", "sphere", "cylinder")
        self.rigid_body = kind
        self._rigid_xyz = torch.as_tensor(rigid_xyz, dtype=torch.float32).to(self.dev).contiguous()
        self._rigid_center = (C.c_float * 3)(*[float(c) for c in center])
        if kind == "cuboid":
            prm = [n * 2.0 * particle_radius / 2.0 for n in cuboid_num]
        elif kind == "sphere":
            prm = [float(sphere_radius), 0.0, 0.0]
        else:
            prm = [float(cylinder_radius), cylinder_num[1] * 2.0 * particle_radius / 2.0, 0.0]
        self._rigid_prm = (C.c_float * 3)(*prm)

    def setup_rigid_body(self, optim_args):
        """The rigid-body block of the reference's setup_constants (gm_dynamics.py:139-157; fields of FD/arguments/__init__.py:423-431)
        followed by create_rigid_body(): reads `rigid_body`, `rigid_body_center` (render units), `rigid_particle_radius`,
        `rigid_cuboid_num`, `rigid_sphere_radius` / `_num`, `rigid_cylinder_radius` / `_num` from `optim_args`, samples the body's
        surface (emitter.py) and registers it for the projections."""
        self.rigid_body = optim_args.rigid_body
        self.rigid_particle_radius = float(optim_args.rigid_particle_radius)
        self.rigid_particle_diameter = 2 * self.rigid_particle_radius
        self.rigid_body_center = torch.tensor(optim_args.rigid_body_center, dtype=torch.float32) * self.scale_factor
        self.rigid_cuboid_num = list(optim_args.rigid_cuboid_num)
        self.rigid_sphere_radius, self.rigid_sphere_num = float(optim_args.rigid_sphere_radius), int(optim_args.rigid_sphere_num)
        self.rigid_cylinder_radius, self.rigid_cylinder_num = float(optim_args.rigid_cylinder_radius), list(optim_args.rigid_cylinder_num)
        self.create_rigid_body()
        self.set_rigid_body(self.rigid_body, self.rigid_body_center.tolist(), self._rigid_xyz, cuboid_num=self.rigid_cuboid_num,
                            particle_radius=self.rigid_particle_radius, sphere_radius=self.rigid_sphere_radius,
                            cylinder_radius=self.rigid_cylinder_radius, cylinder_num=self.rigid_cylinder_num)

    def _rigid_project(self, pts, cap):
        M, N = self._rigid_xyz.size(0), pts.size(0)
        if N == 0 or M == 0:
            return 0
        lib = L.lib()
        grid = torch.empty(lib.fnx_grid_bytes(M), dtype=torch.uint8, device=self.dev)
        n_inside = torch.zeros(1, dtype=torch.int32, device=self.dev)
        kind = {"cuboid": 0, "sphere": 1, "cylinder": 2}[self.rigid_body]
        with torch.cuda.device(self.dev):
            L.check(lib.fnx_rigid_project(grid.data_ptr(), self._rigid_xyz.data_ptr(), M, pts.data_ptr(), N, kind, self._rigid_center,
                                          self._rigid_prm, self.H, cap, n_inside.data_ptr(), self._st()))
        return n_inside

    @torch.no_grad()
    def project_rigid_body_constraints(self):
        """Hidden particles inside the body snap to their nearest rigid sample within H (no neighbour cap)."""
        return self._rigid_project(self._estimate_xyz, 0)

    @torch.no_grad()
    def project_rigid_body_constraints_for_visual_particles(self):
        """Same for the visual particles; the reference's radius() call keeps torch_cluster's default cap of 32."""
        return self._rigid_project(self._visual_xyz, 32)

    # -- per-frame checkpoints in the reference's on-disk layout (gm_fluid.py:1653-1911; fluidnexus_b200/io.py) ------------------
    _SCALAR_ATTRS = dict(secs="_secs", particle_id_max="_particle_id_max")   # scalar key -> attribute, where the names differ

    def _scalar_values(self):
        """The dictionary save_hidden writes to frame_XXX_scalar_values.json (gm_fluid.py:1694-1716), from the solver's attributes;
        counters this class does not keep (total_*_iterations, remove_out_boundary) are written as 0 / False unless set."""
        from . import io as IO
        fallback = dict(remove_out_boundary=False, emit_counter=0, total_iterations=0, total_sim_iterations=0, total_tb_log_iterations=0,
                        particle_id_max=int(self._xyz.shape[0]), emit_ratio_hidden=self._emit_setting("emit_ratio_hidden"),
                        emit_ratio_visual=self._emit_setting("emit_ratio_visual"))
        out = {}
        for k in IO.SCALAR_KEYS:
            attr = self._SCALAR_ATTRS.get(k, k)
            out[k] = getattr(self, attr) if hasattr(self, attr) else fallback[k]
        return out

    @torch.no_grad()
    def save_hidden(self, checkpoint_path, frame_idx):
        from . import io as IO
        N = self._xyz.shape[0]
        state = {k: getattr(self, "_" + k) for k in ("xyz", "estimate_xyz", "buoyancy", "force", "velocity", "imass", "counts")}
        state["gravity"] = torch.tensor([float(c) for c in self._gravity], dtype=torch.float32).reshape(1, 3)
        state["particle_id"] = getattr(self, "_particle_id", None)
        if state["particle_id"] is None:
            state["particle_id"] = torch.arange(N).unsqueeze(1)
        IO.save_hidden(checkpoint_path, frame_idx, state, self._scalar_values())

    @torch.no_grad()
    def load_hidden(self, checkpoint_path, frame_idx):
        """Restores the hidden-particle state and the scalars the reference restores (gm_fluid.py:1811-1884)."""
        from . import io as IO
        state, scal = IO.load_hidden(checkpoint_path, frame_idx, defaults=dict(emit_ratio_hidden=self._emit_setting("emit_ratio_hidden"),
                                                                              emit_ratio_visual=self._emit_setting("emit_ratio_visual")))
        for k in ("xyz", "estimate_xyz", "buoyancy", "force", "velocity", "imass", "counts"):
            setattr(self, "_" + k, torch.from_numpy(state[k]).to(self.dev).contiguous())
        self._gravity = (C.c_float * 3)(*[float(c) for c in state["gravity"].reshape(-1)])
        self._particle_id = torch.from_numpy(state["particle_id"].astype("int64")).to(self.dev).reshape(-1, 1)
        self.scale_factor, self.buoyancy_decay_rate = float(scal["scale_factor"]), float(scal["buoyancy_decay_rate"])
        self.remove_out_boundary = bool(scal["remove_out_boundary"])
        self._secs, self.alpha, self.k, self.p0 = float(scal["secs"]), float(scal["alpha"]), float(scal["k"]), float(scal["p0"])
        self.buoyancy_max_y, self.min_neighbors = float(scal["buoyancy_max_y"]), int(scal["min_neighbors"])
        self.emit_ratio_hidden, self.emit_ratio_visual = scal["emit_ratio_hidden"], scal["emit_ratio_visual"]
        self.emit_counter, self.total_iterations = int(scal["emit_counter"]), int(scal["total_iterations"])
        self.total_sim_iterations, self.total_tb_log_iterations = int(scal["total_sim_iterations"]), int(scal["total_tb_log_iterations"])
        self._particle_id_max = int(scal["particle_id_max"]) or int(self._xyz.shape[0])
        return scal

    @torch.no_grad()
    def save_visual(self, checkpoint_path, frame_idx, scale=True):
        from . import io as IO
        IO.save_visual(checkpoint_path, frame_idx, {k: getattr(self, "_" + k) for k in IO.VISUAL_ARRAYS}, self.scale_factor, scale=scale)

    @torch.no_grad()
    def load_visual(self, checkpoint_path, frame_idx, scale=True, color_3ch=False):
        from . import io as IO
        for k, a in IO.load_visual(checkpoint_path, frame_idx, self.scale_factor, scale=scale, color_3ch=color_3ch).items():
            setattr(self, "_" + k, torch.from_numpy(a).to(self.dev).contiguous())
        return int(self._visual_xyz.shape[0])

    @torch.no_grad()
    def load_visual_smoothed(self, checkpoint_path, frame_idx, scale=True, window_size=5, smoothed_color=True, smoothed_scales=True,
                             smoothed_rotation=True, smoothed_opacity=True):
        """gm_dynamics.py:2093-2150: the temporally smoothed level-two attributes (files *_smoothed_ws<window>.npy) for future prediction."""
        from . import io as IO
        which = tuple(n for n, on in (("visual_color", smoothed_color), ("visual_scales", smoothed_scales),
                                      ("visual_rotation", smoothed_rotation), ("visual_opacity", smoothed_opacity)) if on)
        for k, a in IO.load_visual(checkpoint_path, frame_idx, self.scale_factor, scale=scale, smoothed_window=window_size, smoothed=which).items():
            setattr(self, "_" + k, torch.from_numpy(a).to(self.dev).contiguous())
        return int(self._visual_xyz.shape[0])

    def save_all(self, checkpoint_path, frame_idx):
        self.save_hidden(checkpoint_path, frame_idx)
        self.save_visual(checkpoint_path, frame_idx)

    def load_all(self, checkpoint_path, frame_idx):
        self.load_hidden(checkpoint_path, frame_idx)
        self.load_visual(checkpoint_path, frame_idx)

    # -- quantity snapshots (gm_dynamics.py:1938-1976), same method names -----------------------------------------------------------
    def _snapshot(self, kind, path, a):
        from . import io as IO
        t = dict(xyz=self._xyz, estimate_xyz=self._estimate_xyz, visual_xyz=self._visual_xyz, rigid_xyz=getattr(self, "_rigid_xyz", None))
        return IO.save_particles(kind, path, t, a, scale_factor=self.scale_factor)

    def save_particles_frame(self, quantities_path, frame_idx):
        return self._snapshot("frame", quantities_path, frame_idx)

    def save_particles_simulation(self, quantities_path, index):
        return self._snapshot("simulation", quantities_path, index)

    def save_particles_simulation_guess(self, quantities_path, index):
        return self._snapshot("simulation_guess", quantities_path, index)

    def save_particles_rigid_body(self, quantities_path, frame_idx):
        return self._snapshot("rigid_body", quantities_path, frame_idx)

    # -- future prediction: the loop of FD/entries_fluid_nexus/future_simulation.py:118-175 ---------------------------------------
    @staticmethod
    def future_p0(p0_recon, p0_future, future_time_index, decay_frames):
        """Rest density of future frame t: decays linearly from the reconstruction's p0 to p0_future over `decay_frames` frames
        (future_simulation.py:120)."""
        return p0_future + (p0_recon - p0_future) * (1 - min(1, future_time_index / decay_frames))

    def predict(self, future_pred_frames, first_frame_index=0, solver_iterations_future=3, p0_future=1.5, decay_frames_future_p0=30,
                wind_since=-1, use_level_two_in_future=False, on_frame=None):
        """Simulates `future_pred_frames` frames past the last reconstructed one, in the reference's order per frame: p0 schedule,
        remove_invalid_particles, (first frame) remove_invisible_bottom_visual_particles, emit_new_particles, guess_hidden_particles
        (wind from frame `wind_since` on), `solver_iterations_future` x project_gas_constraints, confirm_guess_hidden_particles,
        update_visual_particles, prepare_future_visual_particles_for_rendering; then `on_frame(self, future_frame_index)` -- where
        the script renders and saves.  Defaults: FD/arguments/__init__.py:310,330,332,418.  Emitter sites must have been prepared
        (prepare_emitter_points(model_args, is_future=True))."""
        p0_recon = self.p0
        for t in range(int(future_pred_frames)):
            frame = first_frame_index + t
            self.p0 = self.future_p0(p0_recon, p0_future, t, decay_frames_future_p0)
            self.remove_invalid_particles()
            if t == 0:
                self.remove_invisible_bottom_visual_particles()
            self.emit_new_particles()
            self.guess_hidden_particles(use_wind=wind_since >= 0 and frame >= wind_since)
            for _ in range(int(solver_iterations_future)):
                self.project_gas_constraints()
            self.confirm_guess_hidden_particles()
            self.update_visual_particles()
            self.prepare_future_visual_particles_for_rendering(use_level_two_in_future)
            if on_frame is not None:
                on_frame(self, frame)

    # -- one simulation tick as the entries run it ---------------------------------------------------------------
    @torch.no_grad()
    def tick(self, solver_iterations=3, stable=False, use_wind=False, count_first=False):
        """future_simulation.py:135-162 (count_first=False) / train_physical_particle.py:206-216 (count_first=True)."""
        self.guess_hidden_particles(stable=stable, use_wind=use_wind)
        if count_first:
            for _ in range(solver_iterations):
                self.update_solver_counts()
        for _ in range(solver_iterations):
            self.project_gas_constraints()
        self.confirm_guess_hidden_particles()
        self.update_visual_particles()
