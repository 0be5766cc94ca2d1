"""CPU, world_size 2 over gloo: the multi-rank plumbing of the step (item assignment, flat gradient bucket,
one all-reduce, replicated update) without any GPU compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fluidnexus_b200.parallel import FlatBucket, assign_items


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


def _worker(rank, world, port, G, V, N, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        by_frame, phys = assign_items(G, V, world, rank)
        fb = FlatBucket(G, N, "cpu")
        for f in phys:                       # owners initialise their frames' parameters
            fb.param[f] = torch.full((N, 3), float(f + 1))
        fb.broadcast_params_from_owners()
        opt = torch.optim.Adam([fb.param], lr=1e-2, eps=1e-15)
        for step in range(3):
            fb.zero_grad()
            for f, views in by_frame.items():
                p, m, v, g = fb.views(f)
                for vw in views:             # image-loss gradients of this rank's views, already scaled by 1/batch
                    g += _fake_item_grad(f, vw, N) * (1.0 / V) * (step + 1)
            for f in phys:                   # view-independent terms: once per frame
                fb.views(f)[3].add_(_fake_physics_grad(f, N))
            fb.all_reduce()
            fb.param.grad = fb.grad.clone()
            opt.step()
        out[rank] = (fb.param.clone(), fb.grad.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("G", [4, 1])
def test_two_ranks_match_single_process(G):
    V, N = 5, 16
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, G, V, N, out), nprocs=2, join=True)
    # single-process reference of the same 3 steps
    param = torch.stack([torch.full((N, 3), float(f + 1)) for f in range(G)])
    opt = torch.optim.Adam([param], lr=1e-2, eps=1e-15)
    for step in range(3):
        grad = torch.zeros_like(param)
        for f in range(G):
            for vw in range(V):
                grad[f] += _fake_item_grad(f, vw, N) * (1.0 / V) * (step + 1)
            grad[f] += _fake_physics_grad(f, N)
        param.grad = grad
        opt.step()
    for r in range(2):
        p, g = out[r]
        assert torch.allclose(g, grad, atol=1e-6), r
        assert torch.allclose(p, param.detach(), atol=1e-6), r
    assert torch.equal(out[0][0], out[1][0])   # replicated parameters stay bit-identical across ranks


def test_frame_lanes_single_lane_runs_frames_in_order():
    """lanes = 1 is plain sequential execution (no CUDA streams involved): results in frame order, one step object."""
    from fluidnexus_b200.parallel import FrameLanes
    made = []
    lanes = FrameLanes(lambda k: made.append(k) or f"step{k}", 1, "cpu")
    assert lanes.n == 1 and made == [0]
    seen = []
    out = lanes.run([3, 1, 2], lambda step, fr: seen.append((step, fr)) or fr * 10)
    assert out == [30, 10, 20] and seen == [("step0", 3), ("step0", 1), ("step0", 2)]
