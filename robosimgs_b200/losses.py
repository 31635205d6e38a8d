"""Fused photometric loss (the step right after the rasterizer in the
training loop the reference delegates, /root/reference/README.md:75).

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
        if a.dtype != torch.float32 or b.shape != a.shape or n % 4:
            raise ValueError("photometric_loss: float32 tensors of equal shape with numel % 4 == 0 expected")
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
