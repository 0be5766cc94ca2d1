"""CPU: host logic of the torch_cluster / torch_scatter / simple_knn drop-ins (fluidnexus_b200/physics.py) that needs no kernel:
radius_graph's edge bookkeeping around `radius` (flow, self loops, the K+1 rule of torch-cluster 1.6.3's radius_graph.py), the
robust cell-size estimate of distCUDA2 and the error behaviour.  The searches themselves run on the GPU only
(tests/test_physics_gpu.py, tests/test_baseline_sizes_gpu.py)."""
import numpy as np
import pytest
import torch

from fluidnexus_b200 import physics as P
from oracle import pbf_ref as O


@pytest.mark.parametrize("loop", [True, False])
@pytest.mark.parametrize("K", [3, 8, 100])
def test_radius_graph_bookkeeping_around_the_search(monkeypatch, loop, K):
    """With the search itself replaced by the oracle's (index-order scan, strict < r^2, first K hits), the drop-in's radius_graph
    must return the oracle's graph -- and the transposed one for flow='target_to_source'."""
    calls = []

    def fake_radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, num_workers=1, batch_size=None):
        calls.append(max_num_neighbors)
        return O.radius(x, y, r, max_num_neighbors=max_num_neighbors)

    monkeypatch.setattr(P, "radius", fake_radius)
    x = torch.tensor(np.random.default_rng(K).uniform(0, 4, (120, 3)), dtype=torch.float32)
    want = O.radius_graph(x, 1.5, loop=loop, max_num_neighbors=K)
    got = P.radius_graph(x, 1.5, loop=loop, max_num_neighbors=K)
    assert calls == [K if loop else K + 1]                      # a query finds itself first: K other neighbours need K + 1 hits
    assert got.dtype == torch.long and torch.equal(got, want)
    if not loop:
        assert bool((got[0] != got[1]).all())
    t2s = P.radius_graph(x, 1.5, loop=loop, max_num_neighbors=K, flow="target_to_source")
    assert torch.equal(t2s, got.flip(0))
    with pytest.raises(AssertionError):
        P.radius_graph(x, 1.5, flow="sideways")


def test_dropins_refuse_what_libfnx_does_not_do():
    x = torch.rand(10, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.radius(x, x, 0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.distCUDA2(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.scatter_min(torch.rand(5), torch.zeros(5, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        P.radius(x, x, 0.5, batch_x=torch.zeros(10, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        P.scatter_min(torch.rand(5, 2), torch.zeros(5, 2, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.density_ratio(x, torch.ones(10, 1), 2.0, 1.5, 100)


def _spacing(p):
    """median nearest-neighbour distance (what the cell should be comparable to)"""
    from scipy.spatial import cKDTree
    d, _ = cKDTree(p).query(p, k=2)
    return float(np.median(d[:, 1]))


@pytest.mark.parametrize("shape", ["volume", "plane", "tilted_plane_outliers", "line", "outliers"])
def test_knn_cell_tracks_the_point_spacing(shape):
    """distCUDA2's grid cell must stay within a small factor of the true spacing: a cell far below it empties the 27-cell
    neighbourhood (exhaustive O(n^2) fallback), one far above it puts thousands of points into a bucket."""
    rng = np.random.default_rng(5)
    n = 20_000
    if shape == "volume":
        p = rng.uniform(0, 1, (n, 3))
    elif shape == "plane":
        p = np.c_[rng.uniform(0, 1, (n, 2)), np.zeros(n)]
    elif shape == "tilted_plane_outliers":
        p = np.c_[rng.uniform(0, 1, (n, 2)), np.zeros(n)]
        p[:20] = rng.uniform(-500, 500, (20, 3))
    elif shape == "line":
        p = np.c_[rng.uniform(0, 1, n), np.full(n, 0.3), np.full(n, -2.0)]
    else:
        p = rng.uniform(0, 1, (n, 3))
        p[:50] = rng.uniform(-1e4, 1e4, (50, 3))
    cell = P._spacing_cell(torch.tensor(p, dtype=torch.float32))
    s = _spacing(p)
    assert 0.5 * s < cell < 12.0 * s, (shape, cell, s)


def test_knn_cell_of_degenerate_clouds_is_positive():
    assert P._spacing_cell(torch.zeros(100, 3)) > 0
    two = torch.zeros(100, 3)
    two[50:] = 1.0
    assert P._spacing_cell(two) > 0
    assert P._spacing_cell(torch.rand(1, 3)) > 0
