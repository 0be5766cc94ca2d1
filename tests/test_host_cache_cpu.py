"""CPU: the ground-truth cache of the fused steps (fluidnexus_b200/step.py:HostTensorCache) never serves a stale image."""
import gc

import torch

from fluidnexus_b200.step import HostTensorCache


def test_same_tensor_is_uploaded_once_until_it_is_edited_in_place():
    c, uploads = HostTensorCache(), []
    up = lambda t: (uploads.append(1), t.clone())[1]
    a = torch.rand(3, 8, 8)
    d0 = c.get(a, up)
    assert c.get(a, up) is d0 and len(uploads) == 1
    a.mul_(0.5)                                     # in-place edit bumps the version counter
    d1 = c.get(a, up)
    assert len(uploads) == 2 and torch.equal(d1, a) and c.get(a, up) is d1
    b = a.clone()                                   # equal content, another object: its own entry
    assert c.get(b, up) is not d1 and len(uploads) == 3


def test_a_new_tensor_in_a_recycled_allocation_is_not_mistaken_for_the_old_one():
    c = HostTensorCache()
    up = lambda t: t.clone()
    seen = set()
    for k in range(200):                            # the allocator hands the freed block (and often the id) out again
        t = torch.full((3, 16, 16), float(k))
        seen.add((t.data_ptr(), id(t)))
        got = c.get(t, up)
        assert float(got[0, 0, 0]) == float(k), k
        del t, got
        gc.collect()
    assert len(seen) < 200                          # addresses / ids really were recycled in this run


def test_bounded_and_dead_entries_go_first():
    c = HostTensorCache(max_entries=4)
    up = lambda t: t.clone()
    keep = [torch.rand(2) for _ in range(3)]
    for t in keep:
        c.get(t, up)
    for _ in range(10):
        c.get(torch.rand(2), up)                    # temporaries: dead as soon as the call returns
    assert len(c) <= 4
    hits = []
    for t in keep:                                  # the live ones survived the evictions of the dead ones
        c.get(t, lambda x: (hits.append(1), x.clone())[1])
    assert len(hits) == 0
