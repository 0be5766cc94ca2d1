"""Host-side mirror of FD/utils/loss_utils.py (l1_loss, ssim) on top of libfnx's fused image-loss kernels."""
import torch

from . import _lib as L


def _f32c(t):
    t = t.detach()
    return (t if t.dtype == torch.float32 else t.float()).contiguous()


def image_loss_raw(img, gt, w_l1, w_ssim, grey=False, want_grad=True):
    """img, gt: [C,H,W] or [V,C,H,W] CUDA tensors.  Returns (l1_mean [V], ssim_mean [V], dL_dimg or None) where
    dL_dimg is the gradient of sum_v (w_l1*l1_mean[v] + w_ssim*(1-ssim_mean[v]))."""
    if not img.is_cuda:
        raise RuntimeError("libfnx image loss needs CUDA tensors (there is no CPU fallback)")
    x, y = _f32c(img), _f32c(gt)
    if x.dim() == 3:
        x, y = x.unsqueeze(0), y.unsqueeze(0)
    V, C, H, W = x.shape
    dev = x.device
    lib = L.lib()
    with torch.cuda.device(dev):
        scratch = torch.empty(lib.fnx_image_loss_bytes(V, C, H, W), dtype=torch.uint8, device=dev)
        l1 = torch.empty(V, dtype=torch.float32, device=dev)
        ss = torch.empty(V, dtype=torch.float32, device=dev)
        g = torch.empty_like(x) if want_grad else None
        L.check(lib.fnx_image_loss(V, C, H, W, x.data_ptr(), y.data_ptr(), int(bool(grey)), float(w_l1), float(w_ssim),
                                   g.data_ptr() if want_grad else None, l1.data_ptr(), ss.data_ptr(), scratch.data_ptr(),
                                   torch.cuda.current_stream(dev).cuda_stream))
    return l1, ss, g


class _ImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, gt, w_l1, w_ssim, grey):
        l1, ss, g = image_loss_raw(img, gt, w_l1, w_ssim, grey)
        ctx.g = g.view_as(img)
        ctx.mark_non_differentiable(l1, ss)
        return (w_l1 * l1 + w_ssim * (1.0 - ss)).sum(), l1, ss

    @staticmethod
    def backward(ctx, go, _a, _b):
        return ctx.g * go, None, None, None, None


def image_loss(img, gt, lambda_dssim=0.2, lambda_image=1.0, grey=False):
    """(1-lambda_dssim)*lambda_image*l1_loss + lambda_dssim*lambda_image*(1-ssim), summed over views if batched;
    differentiable in `img`.  Returns (loss, l1_mean[V], ssim_mean[V])."""
    return _ImageLoss.apply(img, gt, (1.0 - lambda_dssim) * lambda_image, lambda_dssim * lambda_image, grey)


def l1_loss(network_output, gt):
    """FD/utils/loss_utils.py:9-10."""
    return _ImageLoss.apply(network_output, gt, 1.0, 0.0, False)[0]


def ssim(img1, img2, window_size=11, size_average=True):
    """FD/utils/loss_utils.py:26-35 (window 11, size_average=True only)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("libfnx ssim implements the reference's only configuration: window 11, size_average")
    return 1.0 - _ImageLoss.apply(img1, img2, 0.0, 1.0, False)[0]
