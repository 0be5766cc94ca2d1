"""CPU, world_size 2 over gloo: the multi-rank plumbing of the step (item assignment, flat gradient bucket,
one all-reduce, replicated update) without any GPU compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fluidnexus_b200.parallel import FlatBucket, assign_items, plan_items


def test_assignment_covers_every_item_once():
    for G, V, world in [(8, 5, 1), (8, 5, 2), (8, 5, 4), (8, 5, 8), (1, 5, 2), (2, 5, 8), (3, 5, 2), (1, 5, 8)]:
        seen, phys = [], []
        for r in range(world):
            by_frame, pf = assign_items(G, V, world, r)
            seen += [(f, v) for f, vs in by_frame.items() for v in vs]
            phys += sorted(pf)
        assert sorted(seen) == [(f, v) for f in range(G) for v in range(V)], (G, V, world)
        assert sorted(phys) == list(range(G)), (G, V, world)       # physics of a frame computed by exactly one rank
    # whole frames stay on one rank while frames >= ranks
    by_frame, _ = assign_items(8, 5, 4, 1)
    assert all(vs == [0, 1, 2, 3, 4] for vs in by_frame.values()) and sorted(by_frame) == [1, 5]
    # fewer frames than ranks: views straddle ranks
    assert assign_items(1, 5, 2, 0)[0] == {0: [0, 1, 2]} and assign_items(1, 5, 2, 1)[0] == {0: [3, 4]}


def _fake_item_grad(f, v, n):
    g = torch.Generator().manual_seed(1000 * f + v)
    return torch.randn(n, 3, generator=g)


def _fake_physics_grad(f, n):
    g = torch.Generator().manual_seed(77 + f)
    return torch.randn(n, 3, generator=g)


def _adam(p, g, m, v, step, lr=1e-2, b1=0.9, b2=0.999, eps=1e-15):
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.sub_(lr * (m / (1 - b1 ** step)) / ((v / (1 - b2 ** step)).sqrt() + eps))


def _worker(rank, world, port, G, V, N, out):
    """Mirrors bench.py:one_step: NO zero_grad anywhere -- a rank OVERWRITES the slots of the frames it holds items of (as
    fnx_pbf_combine_grad does), begin_step() clears the shared slots it holds nothing of, frames wholly on one rank are updated
    locally, shared frames after the all-reduce on every rank."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = plan_items(G, V, world, rank)
        fb = FlatBucket(G, N, "cpu", plan)
        for f in plan.physics_frames:        # owners initialise their frames' parameters
            fb.param[f] = torch.full((N, 3), float(f + 1))
        fb.broadcast_params_from_owners()
        lo, hi = fb.shared_range()
        for step in range(1, 4):
            fb.begin_step()
            for f, views in plan.by_frame.items():
                p, m, v, g = fb.views(f)
                part = torch.zeros(N, 3)
                for vw in views:             # image-loss gradients of this rank's views, already scaled by 1/batch
                    part += _fake_item_grad(f, vw, N) * (1.0 / V) * step
                if f in plan.physics_frames:  # view-independent terms: once per frame, on its owner
                    part += _fake_physics_grad(f, N)
                g.copy_(part)                # overwrite, like the fused step
                if f not in plan.shared:
                    _adam(p, g, m, v, step)
                fb.losses[f, 0] = float(len(views))
            fb.all_reduce()
            for f in range(lo, hi):
                _adam(*fb.views(f)[:1], fb.grad[f], fb.exp_avg[f], fb.exp_avg_sq[f], step)
            fb.all_reduce_losses()
        table = fb.reduced_losses()
        out[rank] = (fb.gather_params().clone(), fb.grad.clone(), table.clone(), sorted(plan.shared))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("G,world", [(4, 2), (1, 2), (3, 2), (2, 3)])
def test_ranks_match_single_process(G, world):
    V, N = 5, 16
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, G, V, N, out), nprocs=world, join=True)
    # single-process reference of the same 3 steps
    param = torch.stack([torch.full((N, 3), float(f + 1)) for f in range(G)])
    m, v = torch.zeros_like(param), torch.zeros_like(param)
    for step in range(1, 4):
        grad = torch.zeros_like(param)
        for f in range(G):
            for vw in range(V):
                grad[f] += _fake_item_grad(f, vw, N) * (1.0 / V) * step
            grad[f] += _fake_physics_grad(f, N)
        _adam(param, grad, m, v, step)
    for r in range(world):
        p, g, table, shared = out[r]
        assert torch.allclose(p, param, atol=1e-6), r
        for f in shared:                       # the reduced slots hold the full gradient on every rank
            assert torch.allclose(g[f], grad[f], atol=1e-6), (r, f)
        assert torch.equal(table[:, 0], torch.full((G,), float(V)))   # every (frame, view) item was processed exactly once
    assert torch.equal(out[0][0], out[1][0])


def test_plan_shapes():
    p = plan_items(16, 5, 8, 3)
    assert p.shared == [] and p.local == [3, 11] and p.physics_frames == {3, 11}
    p = plan_items(1, 5, 2, 1)
    assert p.shared == [0] and p.local == [] and p.by_frame == {0: [3, 4]} and p.physics_frames == set() and p.owner(0) == 0
    p = plan_items(2, 5, 3, 1)   # 10 items in blocks of 4: rank 1 holds the tail of frame 0 and the head of frame 1
    assert p.shared == [0, 1] and p.by_frame == {0: [4], 1: [0, 1, 2]} and p.physics_frames == {1}
    p = plan_items(1, 5, 8, 6)   # more ranks than items: this rank idles but still takes part in the collectives
    assert p.by_frame == {} and p.shared == [0]


def test_frame_lanes_single_lane_runs_frames_in_order():
    """lanes = 1 is plain sequential execution (no CUDA streams involved): results in frame order, one step object."""
    from fluidnexus_b200.parallel import FrameLanes
    made = []
    lanes = FrameLanes(lambda k: made.append(k) or f"step{k}", 1, "cpu")
    assert lanes.n == 1 and made == [0]
    seen = []
    out = lanes.run([3, 1, 2], lambda step, fr: seen.append((step, fr)) or fr * 10)
    assert out == [30, 10, 20] and seen == [("step0", 3), ("step0", 1), ("step0", 2)]


def test_a_frame_keeps_its_lane_whatever_subset_is_run():
    """Captured iterations of a frame use its lane's PhysicalStep (side stream, loss scratch): the dealing must not depend on
    which subset of the frames a later run() is handed."""
    from fluidnexus_b200.parallel import FrameLanes
    lanes = FrameLanes(lambda k: f"step{k}", 1, "cpu")
    lanes.n = 3                                       # (several lanes need CUDA streams; the dealing itself does not)
    assert [lanes.lane_of(f) for f in (10, 11, 12, 13, 14)] == [0, 1, 2, 0, 1]
    assert [lanes.lane_of(f) for f in (14, 12, 10)] == [1, 2, 0]          # a subset, in another order: same lanes
    assert lanes.lane_of(99) == 2 and lanes.lane_of([1, 2]) is None       # new frame: next lane; unhashable: positional
