"""GPU: the fused L1+SSIM kernels against the plain-torch fp32 restatement of FD/utils/loss_utils.py
(oracle/pbf_ref.py:l1_loss, ssim, image_loss) and its autograd gradient.  Floating point: values rel 1e-5,
gradients rel-L2 1e-4 (the reference's own conv2d path is fp32 with TF32 disabled)."""
import numpy as np
import pytest
import torch

from fluidnexus_b200 import losses as FL
from oracle import pbf_ref as O

pytestmark = pytest.mark.gpu


def _imgs(C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    base = torch.stack([0.5 + 0.4 * torch.sin(xx / (5.0 + c) + c) * torch.cos(yy / 7.0) for c in range(C)])
    img = (base + 0.05 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    gt = (base.roll(2, 2) + 0.05 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    return img, gt


@pytest.mark.parametrize("C,H,W,grey", [(3, 64, 64, False), (1, 70, 45, False), (3, 50, 83, True), (3, 16, 16, False)])
def test_image_loss_matches_torch(libfnx, C, H, W, grey):
    torch.backends.cudnn.allow_tf32 = False
    prm = O.PBFParams()
    img, gt = _imgs(C, H, W, C * H + W)
    ref_img = img.clone().double().requires_grad_(True)
    ref, l1r, ssr = O.image_loss(prm, ref_img, gt.double(), grey=grey)
    ref.backward()
    x = img.cuda().requires_grad_(True)
    got, l1, ss = FL.image_loss(x, gt.cuda(), prm.lambda_dssim, prm.lambda_image, grey=grey)
    got.backward()
    assert abs(float(l1[0]) - float(l1r)) < 1e-5 * abs(float(l1r))
    assert abs((1 - float(ss[0])) - float(ssr)) < 2e-5
    assert abs(float(got) - float(ref)) < 2e-5
    r = (x.grad.cpu().double() - ref_img.grad).norm() / ref_img.grad.norm()
    assert r < 1e-4, r


def test_batched_views_and_dropin_functions(libfnx):
    img0, gt0 = _imgs(3, 48, 40, 1)
    img1, gt1 = _imgs(3, 48, 40, 2)
    x = torch.stack([img0, img1]).cuda().requires_grad_(True)
    gt = torch.stack([gt0, gt1]).cuda()
    loss, l1, ss = FL.image_loss(x, gt)
    loss.backward()
    for v, (a, b) in enumerate([(img0, gt0), (img1, gt1)]):
        xa = a.cuda().requires_grad_(True)
        lv, l1v, ssv = FL.image_loss(xa, b.cuda())
        lv.backward()
        assert torch.allclose(l1[v], l1v[0]) and torch.allclose(ss[v], ssv[0])
        assert torch.allclose(x.grad[v], xa.grad, atol=1e-9)
    # reference-named functions
    assert abs(float(FL.l1_loss(x[0], gt[0])) - float(O.l1_loss(img0, gt0))) < 1e-6
    assert abs(float(FL.ssim(x[0], gt[0])) - float(O.ssim(img0.unsqueeze(0), gt0.unsqueeze(0)))) < 1e-5
