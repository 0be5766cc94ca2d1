"""The restatements against numbers the REFERENCE ITSELF computed: tests/golden/pyref_python.npz was produced by importing
the reference's own pure-Python modules (utils/loss_utils.py, graphics_utils.py, general_utils.py, sh_utils.py) on the CPU
(tools/make_python_golden.py).  CPU tests pin oracle/pbf_ref.py, the synthetic cameras and the host helpers; the GPU tests
compare libfnx's fused kernels with the same fixtures directly."""
import math
import os

import numpy as np
import pytest
import torch

from fluidnexus_b200 import background as B
from fluidnexus_b200 import io as IO
from fluidnexus_b200 import synthetic as S
from oracle import pbf_ref as O

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pyref_python.npz"))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


@pytest.mark.parametrize("tag", ["rgb", "grey1"])
def test_oracle_image_losses_equal_the_reference(tag):
    img = torch.tensor(Z[f"loss_{tag}_img"], requires_grad=True)
    gt = torch.tensor(Z[f"loss_{tag}_gt"])
    l1, ss = O.l1_loss(img, gt), O.ssim(img, gt)
    (0.8 * l1 + 0.2 * (1.0 - ss)).backward()
    assert abs(l1.item() - float(Z[f"loss_{tag}_l1"])) < 1e-14 and abs(ss.item() - float(Z[f"loss_{tag}_ssim"])) < 1e-12
    assert rel(img.grad.numpy(), Z[f"loss_{tag}_grad"]) < 1e-12


def test_oracle_grey_image_loss_equals_the_reference_entries():
    prm = O.PBFParams()
    img = torch.tensor(Z["loss_greyed_img"], requires_grad=True)
    total, l1, s = O.image_loss(prm, img, torch.tensor(Z["loss_greyed_gt"]), grey=True)
    total.backward()
    assert abs(l1.item() - float(Z["loss_greyed_l1"])) < 1e-14 and abs((1.0 - s.item()) - float(Z["loss_greyed_ssim"])) < 1e-12
    assert rel(img.grad.numpy(), Z["loss_greyed_grad"]) < 1e-12


@pytest.mark.parametrize("thr", [0.004, 0.0005])
def test_oracle_distance_loss_equals_the_reference(thr):
    x = torch.tensor(Z["dist_points"], requires_grad=True)
    v = O.distance_loss(x, thr)
    v.backward()
    assert abs(v.item() - float(Z[f"dist_{thr}_value"])) <= 1e-9 * abs(float(Z[f"dist_{thr}_value"]))
    assert rel(x.grad.numpy(), Z[f"dist_{thr}_grad"]) < 1e-9      # (cdist and the explicit difference round differently)
    p = torch.tensor(Z["dist_points"])
    assert abs(O.l2_loss(p, torch.tensor(Z["dist_points"][::-1].copy())).item() - float(Z["l2_value"])) < 1e-15


def test_synthetic_camera_matrices_equal_the_reference():
    """FD/scene/camera.py:90-110 through graphics_utils.get_world_2_view2 / get_projection_matrix."""
    cam = S.SyntheticCamera(Z["cam_R"], Z["cam_T"], 0.69, 0.55, 64, 48, "golden")
    assert np.array_equal(cam.world_view_transform.numpy(), Z["cam_world_view_transform"])
    assert np.array_equal(cam.projection_matrix.numpy(), Z["cam_projection_matrix"])
    assert np.array_equal(cam.full_proj_transform.numpy(), Z["cam_full_proj_transform"])
    assert np.allclose(cam.camera_center.numpy(), Z["cam_center"], atol=0, rtol=0)
    assert np.array_equal(S.world_to_view(Z["cam_R"], Z["cam_T"], np.array([0.5, -0.25, 0.125]), 2.0), Z["cam_w2v_shifted"])


def test_host_helpers_equal_the_reference():
    f = B.expon_lr(1.6e-4, 1.6e-6, lr_delay_mult=0.01, max_steps=30_000)
    g = B.expon_lr(1e-2, 1e-4, lr_delay_steps=100, lr_delay_mult=0.01, max_steps=1000)
    for s, a, b in zip(Z["lr_steps"], Z["lr_plain"], Z["lr_delayed"]):
        assert f(int(s)) == pytest.approx(float(a), rel=1e-12) and g(int(s)) == pytest.approx(float(b), rel=1e-12)
    assert np.allclose(B.inverse_sigmoid(torch.tensor(Z["invsig_x"])).numpy(), Z["invsig_y"], rtol=1e-6)
    assert np.allclose((Z["sh_rgb"] - 0.5) / IO.C0, Z["sh_dc"], rtol=1e-15)      # the f_dc columns of the background PLY


@pytest.mark.gpu
@pytest.mark.parametrize("tag,grey", [("rgb", False), ("grey1", False), ("greyed", True)])
def test_fused_image_loss_kernels_equal_the_reference(libfnx, tag, grey):
    """libfnx's L1/SSIM kernels (fp32) against the reference's own loss_utils outputs (fp64): values rel 1e-5, gradient
    rel-L2 1e-4."""
    from fluidnexus_b200 import losses as FL
    x = torch.tensor(Z[f"loss_{tag}_img"], dtype=torch.float32).cuda().requires_grad_(True)
    gt = torch.tensor(Z[f"loss_{tag}_gt"], dtype=torch.float32).cuda()
    total, l1, ss = FL.image_loss(x, gt, 0.2, 1.0, grey=grey)
    total.backward()
    assert abs(float(l1[0]) - float(Z[f"loss_{tag}_l1"])) < 1e-5 * float(Z[f"loss_{tag}_l1"])
    assert abs(float(ss[0]) - float(Z[f"loss_{tag}_ssim"])) < 2e-5
    assert rel(x.grad.cpu().numpy(), Z[f"loss_{tag}_grad"]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("thr", [0.004, 0.0005])
def test_pair_distance_kernel_equals_the_reference(libfnx, thr):
    from fluidnexus_b200 import physics as P
    x = torch.tensor(Z["dist_points"], dtype=torch.float32).cuda().requires_grad_(True)
    v = P.pair_distance_loss(x, thr)
    v.backward()
    assert abs(float(v) - float(Z[f"dist_{thr}_value"])) < 2e-4 * abs(float(Z[f"dist_{thr}_value"])) + 1e-12
    assert rel(x.grad.cpu().numpy(), Z[f"dist_{thr}_grad"]) < 2e-4
