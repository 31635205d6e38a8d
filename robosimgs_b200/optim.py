"""Fused Adam for the Gaussian parameters (the step right after the
rasterizer's backward in the reconstruction loop the reference delegates, /root/reference/README.md:75).

``FusedAdam`` mirrors ``torch.optim.Adam`` (param groups with per-group ``lr``, ``betas``, ``eps``;
``step()``, ``zero_grad()``, ``state_dict()``-style state in ``self.state``) for dense fp32 CUDA
parameters and updates every tensor of the model in ONE kernel launch of libb200gs
(b200gs_adam_step) instead of torch's multi-tensor foreach chain.  No weight decay / amsgrad (3DGS
uses neither).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch

from . import _cabi


class FusedAdam:
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        params = list(params)
        if params and not isinstance(params[0], dict):
            params = [{"params": params}]
        self.param_groups = []
        for g in params:
            g = dict(g)
            g["params"] = list(g["params"])
            g.setdefault("lr", lr)
            self.param_groups.append(g)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.state: dict = {}
        self.step_count = 0

    def zero_grad(self, set_to_none: bool = True) -> None:
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self) -> None:
        L = _cabi.lib()
        todo = []
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _cabi.B200GSError("FusedAdam needs contiguous float32 CUDA parameters; there is no CPU fallback")
                st = self.state.get(p)
                if st is None:
                    st = self.state[p] = {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
                todo.append((p, p.grad.contiguous(), st["exp_avg"], st["exp_avg_sq"], float(g["lr"])))
        if not todo:
            return
        self.step_count += 1
        dev = todo[0][0].device
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for i in range(0, len(todo), _cabi.ADAM_MAX_GROUPS):
                chunk = todo[i:i + _cabi.ADAM_MAX_GROUPS]
                arr = (_cabi.B200GSAdamGroup * len(chunk))()
                for k, (p, gr, m, v, lr) in enumerate(chunk):
                    arr[k] = _cabi.B200GSAdamGroup(p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(),
                                                   p.numel(), lr, 0.0)
                _cabi.check(L.b200gs_adam_step(arr, C.c_int32(len(chunk)), C.c_float(self.betas[0]),
                                               C.c_float(self.betas[1]), C.c_float(self.eps),
                                               C.c_int32(self.step_count), stream))
