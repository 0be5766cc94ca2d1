"""Seeded scenes shared by the oracle tests, the GPU parity tests and tools/make_golden.py."""
import numpy as np

from fluidnexus_b200 import synthetic as S

# name -> (builder, channels, image (W,H), camera index, background)
SCENES = {}


def scene(name):
    def deco(fn):
        SCENES[name] = fn
        return fn
    return deco


@scene("mixed_ch3_96")
def _mixed_ch3():
    gs = S.cat_sets(S.fluid_gaussians(1500, 3, seed=10), S.background_gaussians(2500, 3, seed=11))
    cam = S.make_cameras(5, 96, height=80)[1]
    return gs, cam, np.array([0.1, 0.2, 0.3], np.float32)


@scene("fluid_ch1_112")
def _fluid_ch1():
    gs = S.fluid_gaussians(4000, 1, seed=20)
    cam = S.make_cameras(5, 112)[3]
    return gs, cam, np.array([0.0], np.float32)


@scene("ragged_ch3_70x45")
def _ragged():
    # image size not a multiple of 16, big splats, high opacity -> early termination paths
    gs = S.random_gaussians(1200, 3, seed=30, spread=0.2, log_scale=(-4.0, -2.5))
    cam = S.make_cameras(5, 70, height=45)[2]
    return gs, cam, np.array([1.0, 1.0, 1.0], np.float32)


@scene("dense_ch1_64")
def _dense():
    gs = S.random_gaussians(3000, 1, seed=40, spread=0.08, log_scale=(-4.5, -3.0))
    cam = S.make_cameras(5, 64)[0]
    return gs, cam, np.array([0.5], np.float32)


@scene("behind_ch3_48")
def _behind():
    # half of the points are behind / too close to the camera (z <= 0.2 cull), some far outside the frustum
    gs = S.random_gaussians(800, 3, seed=50, spread=1.4, log_scale=(-4.0, -2.0))
    cam = S.make_cameras(5, 48)[4]
    return gs, cam, np.array([0.0, 0.0, 0.0], np.float32)


def build(name):
    gs, cam, bg = SCENES[name]()
    return gs, cam, bg, S.raster_inputs(gs, cam, bg)


def dL_dpix(name, shape):
    seed = sum(ord(c) for c in name)
    return np.random.default_rng(seed).normal(size=shape).astype(np.float32)
