"""Data-parallel plumbing of the training step.

Work items are (frame, view) pairs.  What has to cross ranks follows from where a frame's views live:

* a frame whose views all sit on ONE rank (frames >= ranks: whole frames are dealt round-robin) needs no gradient
  exchange at all -- the reference's in-process gradient cache (cache_gradient_current / set_batch_gradient_current,
  FD/gaussian_splatting/gm_fluid.py:419-430: sum over the views, times 1/batch) is complete on that rank, which also
  applies the Adam update locally.  Summing such a frame's slot with the zeros of the other ranks would be an all-gather
  in disguise, and -- unless every rank zeroes the slots it does not own every step -- wrong (the previous step's
  reduced value would be summed again).  Only the per-frame loss table is all-reduced (one tiny collective per step) so
  that rank 0 can log the whole job's losses;
* a frame whose views STRADDLE ranks (fewer frames than ranks: the view loop of one frame is the only parallelism the
  time-sequential algorithm has, train_physical_particle.py:308-379) is the real exchange step: every rank holding
  items of the frame writes its partial gradient (image terms of its views, already scaled by 1/batch; the
  view-independent physics terms on the frame's owner only), ranks without items contribute zeros (their slot is
  cleared every step), ONE all-reduce(sum) over those frames' slots reproduces the gradient cache, and every rank
  applies the same Adam update to its replica.

Parameters, Adam moments and gradients of all frames in flight live in flat [G, N, 3] buffers (`FlatBucket`); per-frame
state aliases into them.  The collective is torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from dataclasses import dataclass, field

import torch
import torch.distributed as dist


def assign_items(n_frames, n_views, world, rank):
    """Items (frame, view) owned by `rank`.  Whole frames are dealt round-robin while there are at least as many
    frames as ranks (no view of a frame then crosses a rank); otherwise the flattened item list is split in
    contiguous blocks, so a frame's views may straddle ranks and the all-reduce sums their partial gradients.
    Returns {frame: [views]} and the set of frames whose view-independent physics terms this rank computes
    (exactly one rank per frame: the owner of the frame's first item)."""
    p = plan_items(n_frames, n_views, world, rank)
    return p.by_frame, p.physics_frames


@dataclass
class ShardPlan:
    """Where the (frame, view) items of one step live.  `shared` / `ranks_of` are the same on every rank."""
    n_frames: int
    n_views: int
    world: int
    rank: int
    by_frame: dict = field(default_factory=dict)        # this rank: frame -> [views]
    physics_frames: set = field(default_factory=set)    # this rank: frames whose view-independent terms it computes
    ranks_of: dict = field(default_factory=dict)        # frame -> sorted ranks holding at least one of its views
    shared: list = field(default_factory=list)          # frames whose views straddle ranks (gradient all-reduce + replicated Adam)

    @property
    def local(self):
        """Frames wholly on this rank: gradient complete locally, Adam applied locally (inside the captured iteration)."""
        return sorted(f for f in self.by_frame if f not in self._shared_set)

    @property
    def _shared_set(self):
        return set(self.shared)

    def owner(self, f):
        return self.ranks_of[f][0]


def plan_items(n_frames, n_views, world, rank):
    items = [(f, v) for f in range(n_frames) for v in range(n_views)]
    if n_frames >= world:
        where = {(f, v): f % world for (f, v) in items}
    else:
        per = (len(items) + world - 1) // world
        where = {it: k // per for k, it in enumerate(items)}
    plan = ShardPlan(n_frames, n_views, world, rank)
    for (f, v), r in where.items():
        plan.ranks_of.setdefault(f, set()).add(r)
        if r == rank:
            plan.by_frame.setdefault(f, []).append(v)
    plan.ranks_of = {f: sorted(rs) for f, rs in plan.ranks_of.items()}
    straddle = sorted(f for f, rs in plan.ranks_of.items() if len(rs) > 1)
    # one contiguous slot range for the collective: frames that happen to lie between two straddling frames are treated as
    # shared too (their other ranks contribute zeros)
    plan.shared = list(range(straddle[0], straddle[-1] + 1)) if straddle else []
    plan.physics_frames = {f for f, rs in plan.ranks_of.items() if rs[0] == rank}
    return plan


def _dist_on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class FlatBucket:
    """Flat [G, N, 3] parameter / Adam-moment / gradient buffers; per-frame views alias into them.  `losses` [G, L] is
    the per-frame loss table that is all-reduced for logging."""

    N_LOSS = 16   # = step.LOSS_ROW: [gas, next_gas, exyz, dist | l1 x5 | ssim x5 | pad]

    def __init__(self, n_frames, n_particles, device, plan: ShardPlan = None):
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)
        G, N = n_frames, n_particles
        self.param, self.exp_avg, self.exp_avg_sq, self.grad = z(G, N, 3), z(G, N, 3), z(G, N, 3), z(G, N, 3)
        self.losses = z(G, self.N_LOSS)
        # the reduced loss table is double-buffered: the all-reduce of step k runs on NCCL's own stream while step k+1 is being
        # queued, and is only waited for when somebody reads it (reduced_losses) or its buffer comes round again
        self._losses_global = [z(G, self.N_LOSS), z(G, self.N_LOSS)]
        self._loss_work = [None, None]
        self._loss_slot = 0
        self.plan = plan
        self.step = 0

    def views(self, f):
        return self.param[f], self.exp_avg[f], self.exp_avg_sq[f], self.grad[f]

    def zero_grad(self):
        self.grad.zero_()

    # -- the per-step exchange ---------------------------------------------------------------------------------------
    def shared_range(self):
        """The shared frames as one contiguous slot range [lo, hi) (plan_items makes the set contiguous; empty when every
        frame lives on one rank)."""
        s = self.plan.shared if self.plan is not None else []
        return (s[0], s[-1] + 1) if s else (0, 0)

    def begin_step(self):
        """Clear what this rank will NOT overwrite but the all-reduce will sum: the slots of shared frames in which it holds
        no item.  (Slots it does write are fully overwritten by
        fnx_pbf_combine_grad; slots of frames that live wholly on other ranks are never read.)"""
        lo, hi = self.shared_range()
        for f in range(lo, hi):
            if self.plan is None or f not in self.plan.by_frame:
                self.grad[f].zero_()

    def all_reduce(self):
        """Sum the partial gradients of the frames whose views straddle ranks.  No-op when every frame lives on one rank.
        Without a plan (legacy callers): the whole bucket, which then must have been zeroed with zero_grad()."""
        if not _dist_on():
            return
        if self.plan is None:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
            return
        lo, hi = self.shared_range()
        if hi > lo:
            dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM)

    def all_reduce_losses(self):
        """Per-frame loss rows: every row is written by exactly one rank (the frame's owner adds the view-independent
        terms, image terms are partial sums per rank), so the sum over ranks is the job's loss table.  The collective is
        ASYNCHRONOUS (it only feeds logging): it is ordered after the step's kernels but the compute stream does not wait for it, so
        the ranks are not forced into lock-step every iteration.  Returns the buffer the result lands in; reduced_losses() waits."""
        k = self._loss_slot = self._loss_slot ^ 1
        if self._loss_work[k] is not None:           # this buffer's previous reduction (two steps ago) must have landed
            self._loss_work[k].wait()
            self._loss_work[k] = None
        buf = self._losses_global[k]
        buf.copy_(self.losses)
        if _dist_on():
            self._loss_work[k] = dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True)
        return buf

    def reduced_losses(self, previous=False):
        """The job's loss table of the LAST all_reduce_losses() call (previous=True: of the one before, whose collective has had a
        whole step to finish), safe to read on the current stream."""
        k = self._loss_slot ^ 1 if previous else self._loss_slot
        if self._loss_work[k] is not None:
            self._loss_work[k].wait()                 # makes the current stream wait for the collective (no host block)
            self._loss_work[k] = None
        return self._losses_global[k]

    @property
    def losses_global(self):
        return self.reduced_losses()

    def broadcast_params_from_owners(self):
        """After each rank initialised only the slots of the frames it owns (others zero): sum = every slot from its owner."""
        if _dist_on():
            dist.all_reduce(self.param, op=dist.ReduceOp.SUM)

    def gather_params(self):
        """Every frame's current parameters on every rank (verification only).  Frames that live on one rank are taken from
        that rank; replicated (shared) frames from their owner."""
        out = torch.zeros_like(self.param)
        plan = self.plan
        for f in range(self.param.shape[0]):
            if plan is None or plan.owner(f) == plan.rank:
                out[f].copy_(self.param[f])
        if _dist_on():
            dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out


class FrameLanes:
    """Runs the iterations of independent frames on `lanes` CUDA streams of one GPU.

    One iteration is a chain of dependent launches, several of them far too small to fill 148 SMs (grid builds,
    per-tile scans) and the large ones end in a tail of a few long tiles; with two frames in flight on two streams the
    tails and small kernels of one frame run next to the kernels of the other.  This is the single-GPU form of the
    frame sharding of SURVEY.md 8(e): frames must be independent (separate FrameState, gradients into separate slots).
    `make_step(lane)` builds one PhysicalStep per lane (each owns its side stream, copy stream and loss scratch)."""

    def __init__(self, make_step, lanes, device):
        self.dev = torch.device(device)
        self.n = max(1, int(lanes))
        self.steps = [make_step(k) for k in range(self.n)]
        # (one lane needs no streams or events: the frames simply run in order on the caller's stream)
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.n)] if self.n > 1 else [None]
        self._fork = torch.cuda.Event() if self.n > 1 else None
        self._join = [torch.cuda.Event() for _ in range(self.n)] if self.n > 1 else []
        self._lane_of = {}   # frame -> lane, fixed at first sight

    def lane_of(self, frame):
        """A frame keeps the lane it was first dealt to (round-robin in order of first appearance), whatever subset of the frames
        a later run() is given: its captured iterations use that lane's PhysicalStep (side stream, loss scratch), and two frames
        that share a scratch must never run concurrently."""
        try:
            lane = self._lane_of.get(frame)
            if lane is None:
                lane = self._lane_of[frame] = len(self._lane_of) % self.n
            return lane
        except TypeError:        # unhashable frame objects: positional dealing is the caller's responsibility
            return None

    def run(self, frames, call):
        """call(step, frame) for every frame on the frame's lane (lane_of); everything is ordered after the work already
        queued on the current stream, and the current stream waits for all lanes before this returns.  Returns the
        list of results in frame order."""
        if self.n == 1:
            return [call(self.steps[0], fr) for fr in frames]
        main = torch.cuda.current_stream(self.dev)
        self._fork.record(main)
        used, out = set(), []
        for k, fr in enumerate(frames):
            lane = self.lane_of(fr)
            if lane is None:
                lane = k % self.n
            st = self.streams[lane]
            if lane not in used:
                st.wait_event(self._fork)
                used.add(lane)
            with torch.cuda.stream(st):
                out.append(call(self.steps[lane], fr))
        for lane in sorted(used):
            self._join[lane].record(self.streams[lane])
            main.wait_event(self._join[lane])
        return out
