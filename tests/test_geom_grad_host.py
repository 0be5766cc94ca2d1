"""CPU: the chain rule of geom_bwd_kernel (fluidnexus_b200/csrc/geom_grad.cuh: screen covariance / screen mean / 3-D covariance
backward, written in matrix form) compiled for the HOST with g++ and held to

1. the fp64 rasterizer oracle (oracle/raster_ref.c, the literal restatement of R3/cuda_rasterizer/backward.cu:137-381, pinned to
   the compiled reference by tests/test_oracle_raster.py) on the five seeded scenes, fed with the oracle's own per-Gaussian
   dL/dconic and dL/dmeans2D;
2. finite differences / numpy closed forms of each identity.

The header is the code the CUDA kernel runs (same functions, __host__ __device__); the harness in tests/host/ is test
infrastructure and is never loaded by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import scenes
from oracle.raster_oracle import RasterOracle

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "fluidnexus_b200", "csrc")


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("geomgrad") / "libgeomgrad_host.so")
    # -ffp-contract=off: plain mul/add roundings; the GPU build contracts into FMAs, both sit far inside the tolerances
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", "-I", CSRC,
                    os.path.join(HERE, "host", "geom_grad_host.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    lib.fnx_host_geom_backward.restype = None
    lib.fnx_host_geom_backward.argtypes = ([C.c_int, C.c_void_p, C.c_void_p, C.c_float] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_float, C.c_float]
                                           + [C.c_void_p] * 7)
    lib.fnx_host_screen_cov_grad.argtypes = [C.c_float] * 3 + [C.c_void_p] * 2
    lib.fnx_host_cov3d_backward.argtypes = [C.c_void_p] * 5
    lib.fnx_host_quat_rotation.argtypes = [C.c_void_p] * 2
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, np.float32))


def _rot(q):
    r, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                     [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                     [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])


def _cov6(s, q):
    Q = _rot(q)
    S = Q @ np.diag(np.asarray(s, float) ** 2) @ Q.T
    return np.array([S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]])


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, float) - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("name", sorted(scenes.SCENES))
def test_host_build_of_the_kernel_math_matches_the_oracle(oracle_built, host, name):
    gs, cam, bg, inp = scenes.build(name)
    o = RasterOracle("f64")
    out = o.forward(**inp)
    ref = o.backward(scenes.dL_dpix(name, out["color"].shape))
    P = inp["means3D"].shape[0]
    means, sc, rot = _f32(inp["means3D"]), _f32(inp["scales"]), _f32(inp["rotations"])
    cov = _f32(np.stack([_cov6(inp["scale_modifier"] * sc[i].astype(float), rot[i].astype(float)) for i in range(P)]))
    view, proj = _f32(inp["view"]).reshape(16), _f32(inp["proj"]).reshape(16)
    radii = np.ascontiguousarray(out["radii"], dtype=np.int32)
    g2, gc = _f32(ref["means2D"]), _f32(ref["conic"])
    dm, dc, ds, dq = (np.zeros((P, k), np.float32) for k in (3, 6, 3, 4))
    host.fnx_host_geom_backward(P, _p(means), _p(sc), C.c_float(inp["scale_modifier"]), _p(rot), _p(cov), _p(view), _p(proj), inp["W"], inp["H"],
                                C.c_float(inp["tan_fov_x"]), C.c_float(inp["tan_fov_y"]), _p(radii), _p(g2), _p(gc), _p(dm), _p(dc), _p(ds), _p(dq))
    assert int((radii > 0).sum()) > 50
    # fp32 evaluation of a well-conditioned closed form against the fp64 one
    assert _rel(dc, ref["cov3D"]) < 2e-5
    assert _rel(dm, ref["means3D"]) < 2e-5
    assert _rel(ds, ref["scales"]) < 2e-5
    assert _rel(dq, ref["rotations"]) < 2e-5
    # culled Gaussians get exact zeros
    assert not dm[radii <= 0].any() and not dq[radii <= 0].any()


def test_screen_cov_grad_is_the_derivative_of_the_inverse(host):
    rng = np.random.default_rng(0)
    for _ in range(50):
        a, c = rng.uniform(0.5, 30.0, 2)
        b = rng.uniform(-0.9, 0.9) * np.sqrt(a * c)
        g = rng.normal(size=3)
        F = np.array([[g[0], g[1]], [g[1], g[2]]])
        K = np.linalg.inv(np.array([[a, b], [b, c]]))
        want = -K @ F @ K   # dL/dS2 for L = <F, S2^-1>
        D = np.zeros(3, np.float32)
        host.fnx_host_screen_cov_grad(C.c_float(a), C.c_float(b), C.c_float(c), _p(_f32(g)), _p(D))
        assert np.allclose(D, [want[0, 0], want[0, 1], want[1, 1]], rtol=2e-4, atol=1e-6 * np.abs(want).max())
    # det^2 overflows -> the reference drops the term (1 / inf == 0): zeros, not NaN
    D = np.ones(3, np.float32)
    host.fnx_host_screen_cov_grad(C.c_float(1e20), C.c_float(0.0), C.c_float(1e20), _p(_f32([1, 1, 1])), _p(D))
    assert not D.any()


def test_cov3d_backward_matches_finite_differences(host):
    rng = np.random.default_rng(1)
    for _ in range(20):
        s = rng.uniform(0.01, 0.3, 3)
        q = rng.normal(size=4)          # not normalised on purpose: the reference uses the quaternion as given
        w = rng.normal(size=6)          # L = w . cov6(s, q)
        L = lambda s_, q_: float(w @ _cov6(s_, q_))
        ds, dq = np.zeros(3, np.float32), np.zeros(4, np.float32)
        host.fnx_host_cov3d_backward(_p(_f32(s)), _p(_f32(q)), _p(_f32(w)), _p(ds), _p(dq))
        h = 1e-6
        fd_s = np.array([(L(s + h * e, q) - L(s - h * e, q)) / (2 * h) for e in np.eye(3)])
        fd_q = np.array([(L(s, q + h * e) - L(s, q - h * e)) / (2 * h) for e in np.eye(4)])
        assert np.allclose(ds, fd_s, rtol=2e-4, atol=2e-5 * np.abs(fd_s).max())
        assert np.allclose(dq, fd_q, rtol=2e-4, atol=2e-5 * np.abs(fd_q).max())
    Q = np.zeros(9, np.float32)
    q = _f32([0.3, -0.5, 0.7, 0.2])
    host.fnx_host_quat_rotation(_p(q), _p(Q))
    assert np.allclose(Q.reshape(3, 3), _rot(q.astype(float)), atol=1e-6)
