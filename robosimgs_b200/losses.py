"""Fused photometric loss (the step right after the rasterizer in the
training loop the reference delegates, /root/reference/README.md:75).

``ssim(img, target)`` is the fused SSIM every public 3DGS trainer uses (11x11 Gaussian window, sigma 1.5)
and ``gs_loss(img, target, lambda_dssim)`` = (1 - l) L1 + l (1 - SSIM), the 3DGS training loss.

``photometric_loss(img, target, w_l2, w_l1)`` = mean(w_l2 (img-target)^2 + w_l1 |img-target|) with one
CUDA pass forward and one backward (libb200gs), instead of the ~8 elementwise/reduction launches the
same expression costs in eager PyTorch.  ``mse_loss`` is the fixed cheap loss of the BASELINE metric.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi


class _PhotometricLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, target, w_l2, w_l1):
        L = _cabi.lib()
        if img.device.type != "cuda":
            raise _cabi.B200GSError("b200gs needs CUDA tensors; there is no CPU fallback")
        a = img.contiguous()
        b = target.detach().to(torch.float32).contiguous()
        n = a.numel()
        if a.dtype != torch.float32 or b.shape != a.shape:
            raise ValueError("photometric_loss: float32 tensors of equal shape expected")
        out = torch.empty((), dtype=torch.float32, device=a.device)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(a.device):
            _cabi.check(L.b200gs_photometric_loss(p(a), p(b), C.c_int64(n), C.c_float(w_l2), C.c_float(w_l1), p(out),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        ctx.save_for_backward(a, b)
        ctx.w = (float(w_l2), float(w_l1))
        return out / n

    @staticmethod
    def backward(ctx, grad_out):
        L = _cabi.lib()
        a, b = ctx.saved_tensors
        n = a.numel()
        g = torch.empty_like(a)
        up = grad_out.to(torch.float32).contiguous()
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(a.device):
            _cabi.check(L.b200gs_photometric_loss_backward(
                p(a), p(b), C.c_int64(n), C.c_float(ctx.w[0]), C.c_float(ctx.w[1]), C.c_float(1.0 / n), p(up), p(g),
                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return g, None, None, None


def photometric_loss(img: torch.Tensor, target: torch.Tensor, w_l2: float = 0.0, w_l1: float = 1.0) -> torch.Tensor:
    return _PhotometricLoss.apply(img, target, float(w_l2), float(w_l1))


def mse_loss(img: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean((img - target)^2), fused."""
    return _PhotometricLoss.apply(img, target, 1.0, 0.0)


class _FusedSSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, target):
        L = _cabi.lib()
        if img.device.type != "cuda":
            raise _cabi.B200GSError("b200gs needs CUDA tensors; there is no CPU fallback")
        a = img.contiguous()
        b = target.detach().to(torch.float32).contiguous()
        if a.dtype != torch.float32 or b.shape != a.shape or a.dim() != 3:
            raise ValueError("ssim: float32 [C,H,W] tensors of equal shape expected")
        Cc, H, W = a.shape
        need = ctx.needs_input_grad[0]
        maps = torch.empty((3, Cc, H, W), dtype=torch.float32, device=a.device) if need else None
        out = torch.empty((), dtype=torch.float32, device=a.device)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        with torch.cuda.device(a.device):
            _cabi.check(L.b200gs_ssim_forward(p(a), p(b), C.c_int32(Cc), C.c_int32(H), C.c_int32(W), p(maps), p(out),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        if need:
            ctx.save_for_backward(a, b, maps)
        return out / a.numel()

    @staticmethod
    def backward(ctx, grad_out):
        L = _cabi.lib()
        a, b, maps = ctx.saved_tensors
        Cc, H, W = a.shape
        g = torch.empty_like(a)
        up = grad_out.to(torch.float32).contiguous()
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(a.device):
            _cabi.check(L.b200gs_ssim_backward(p(a), p(b), p(maps), C.c_int32(Cc), C.c_int32(H), C.c_int32(W),
                                               C.c_float(1.0 / a.numel()), p(up), p(g),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return g, None


def ssim(img: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Mean SSIM of two [C,H,W] images (window 11, sigma 1.5, zero padding), fused forward + backward."""
    return _FusedSSIM.apply(img, target)


def gs_loss(img: torch.Tensor, target: torch.Tensor, lambda_dssim: float = 0.2) -> torch.Tensor:
    """(1 - lambda) * L1 + lambda * (1 - SSIM): the 3DGS reconstruction loss, three fused launches forward."""
    return (1.0 - lambda_dssim) * photometric_loss(img, target, 0.0, 1.0) + lambda_dssim * (1.0 - ssim(img, target))
