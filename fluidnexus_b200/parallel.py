"""Data-parallel plumbing of the training step: work items are (frame, view) pairs, the parameters of
all frames in flight are replicated on every rank in flat buffers, every rank fills the gradient-bucket slots of the
items it processed, ONE all-reduce(sum) per step over the flat bucket reproduces the reference's in-process gradient
cache (cache_gradient_current / set_batch_gradient_current, FD/gaussian_splatting/gm_fluid.py:419-430: sum over the
views, times 1/batch -- the 1/batch is folded into the image-loss weights), then one fused Adam launch updates every
frame on every rank.  The collective is torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def assign_items(n_frames, n_views, world, rank):
    """Items (frame, view) owned by `rank`.  Whole frames are dealt round-robin while there are at least as many
    frames as ranks (no view of a frame then crosses a rank); otherwise the flattened item list is split in
    contiguous blocks, so a frame's views may straddle ranks and the all-reduce sums their partial gradients.
    Returns {frame: [views]} and the set of frames whose view-independent physics terms this rank computes
    (exactly one rank per frame: the owner of the frame's first item)."""
    items = [(f, v) for f in range(n_frames) for v in range(n_views)]
    if n_frames >= world:
        mine = [(f, v) for (f, v) in items if f % world == rank]
        owner_of = {f: f % world for f in range(n_frames)}
    else:
        per = (len(items) + world - 1) // world
        mine = items[rank * per:(rank + 1) * per]
        owner_of = {}
        for r in range(world):
            for (f, v) in items[r * per:(r + 1) * per]:
                owner_of.setdefault(f, r)
    by_frame = {}
    for f, v in mine:
        by_frame.setdefault(f, []).append(v)
    physics_frames = {f for f, r in owner_of.items() if r == rank}
    return by_frame, physics_frames


class FlatBucket:
    """Flat, replicated [G, N, 3] parameter / Adam-moment / gradient buffers; per-frame views alias into them."""

    def __init__(self, n_frames, n_particles, device):
        z = lambda: torch.zeros((n_frames, n_particles, 3), dtype=torch.float32, device=device)
        self.param, self.exp_avg, self.exp_avg_sq, self.grad = z(), z(), z(), z()
        self.step = 0

    def views(self, f):
        return self.param[f], self.exp_avg[f], self.exp_avg_sq[f], self.grad[f]

    def zero_grad(self):
        self.grad.zero_()

    def all_reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)

    def broadcast_params_from_owners(self):
        """After each rank initialised only its own frames' slots (others zero): sum = every slot from its owner."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.param, op=dist.ReduceOp.SUM)


class FrameLanes:
    """Runs the iterations of independent frames on `lanes` CUDA streams of one GPU.

    One iteration is a chain of ~45 dependent launches, several of them far too small to fill 148 SMs (grid builds,
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

    def run(self, frames, call):
        """call(step, frame) for every frame, frame k on lane k % lanes; everything is ordered after the work already
        queued on the current stream, and the current stream waits for all lanes before this returns.  Returns the
        list of results in frame order."""
        if self.n == 1:
            return [call(self.steps[0], fr) for fr in frames]
        main = torch.cuda.current_stream(self.dev)
        self._fork.record(main)
        used, out = set(), []
        for k, fr in enumerate(frames):
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
