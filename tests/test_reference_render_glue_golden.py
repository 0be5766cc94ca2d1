"""CPU: the render glue mirrors (fluidnexus_b200/renderer.py) against the reference's OWN glue.  tools/make_render_glue_golden.py
ran renderer/pipe_fluid.py, pipe_dynamics.py and pipe_background.py from /root/reference on the CPU with a recording stand-in
for the rasterizer classes; the mirrors, given the same model, camera and stand-in, must hand the rasterizer the same tensors
and settings and return the same dictionary."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from fluidnexus_b200 import renderer as RD  # noqa: E402
from fluidnexus_b200 import synthetic as S  # noqa: E402
from make_render_glue_golden import CASES, FakeModel, recording_rasterizer, run_case  # noqa: E402  (pure helpers: no reference import)

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pyref_render_glue.npz"))


@pytest.mark.parametrize("name", list(CASES))
def test_mirror_hands_the_rasterizer_what_the_reference_glue_does(name):
    fn_name, kw = CASES[name]
    gm, cam = FakeModel(), S.make_cameras(5, 32, height=24)[1]
    rec = run_case(getattr(RD, fn_name), gm, cam, kw)
    for k, v in rec.items():
        ref = Z[f"{name}__{k}"]
        if isinstance(v, np.ndarray) and v.dtype.kind in "fc":
            assert v.shape == ref.shape and np.array_equal(v, ref), (name, k)
        else:
            assert np.array_equal(np.asarray(v), ref), (name, k, v, ref)


def test_synthetic_camera_equals_the_reference_camera_class():
    """G4: the attributes the pipes and entries read, against an instance of the reference's own scene/camera.py:Camera."""
    c = S.SyntheticCamera(Z["camera__R"], Z["camera__T"], float(Z["camera__fov"][0]), float(Z["camera__fov"][1]), int(Z["camera__hw"][1]),
                          int(Z["camera__hw"][0]), "view_00", timestamp=float(Z["camera__timestamp"]))
    for k in ("world_view_transform", "projection_matrix", "full_proj_transform", "camera_center"):
        assert np.array_equal(getattr(c, k).numpy(), Z[f"camera__{k}"]), k
    assert (c.z_near, c.z_far) == tuple(Z["camera__znear_zfar"]) and (c.image_height, c.image_width) == tuple(Z["camera__hw"])
    # the ground truth the entries upload each iteration is the clamped input image (camera.py:57-70)
    assert np.array_equal(np.clip(Z["camera__image_in"], 0.0, 1.0), Z["camera__original_image"])
