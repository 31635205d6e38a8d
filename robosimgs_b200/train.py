"""Camera-batch data-parallel training step (SURVEY.md 8(e) "Training", 8(f) row 4).

The scene is replicated (robosimgs_b200.sweep.replicate_scene); at every step each rank renders ITS
cameras of the batch through GaussianRasterizer (forward + backward), the per-rank gradients are
averaged with ONE collective over a contiguous bucket (NCCL over NVLink 5 on GPUs -- 236 MB at 1 M
Gaussians; gloo in the CPU tests), and every rank applies the same optimizer step, so the replicas
stay bit-identical without ever broadcasting parameters again.  There is no data-path collective
inside a frame (frames are independent); this is the only exchange step of the training path.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Sequence

import torch
import torch.distributed as dist


class GradBucket:
    """Contiguous fp32 bucket holding the gradients of a fixed list of parameter tensors.

    ``pack()`` copies the ``.grad`` of every parameter into its slice (zeros where a parameter has no
    gradient), ``all_reduce()`` averages the bucket over the process group in one collective,
    ``unpack()`` points every ``.grad`` at its slice of the bucket (views, no copy back)."""

    def __init__(self, params: Iterable[torch.Tensor]):
        self.params: List[torch.Tensor] = list(params)
        if not self.params:
            raise ValueError("GradBucket needs at least one parameter")
        dev, dt = self.params[0].device, self.params[0].dtype
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), dtype=dt, device=dev)
        self.views, o = [], 0
        for p, n in zip(self.params, sizes):
            self.views.append(self.flat[o:o + n].view_as(p))
            o += n

    def pack(self) -> None:
        with torch.no_grad():
            for p, v in zip(self.params, self.views):
                if p.grad is None:
                    v.zero_()
                elif p.grad.data_ptr() != v.data_ptr():
                    v.copy_(p.grad)

    def all_reduce(self, average: bool = True) -> None:
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat)
            if average:
                self.flat.div_(dist.get_world_size())

    def unpack(self) -> None:
        for p, v in zip(self.params, self.views):
            p.grad = v


def backward_or_retry(loss_fn: Callable[..., torch.Tensor], *args, retries: int = 2) -> torch.Tensor:
    """loss_fn(*args).backward() for a loss rendered through a GaussianRasterizer with ``defer_pair_check`` on: the
    host never waits inside the step; if the frame turns out to have been incomplete (PairCapacityExceeded, raised by
    the rasterizer's backward BEFORE any gradient reaches a leaf) the camera is simply rendered again -- the pair
    capacity hint has been raised by then.  Returns the loss."""
    from .rasterizer import PairCapacityExceeded
    for attempt in range(retries + 1):
        loss = loss_fn(*args)
        try:
            loss.backward()
            return loss
        except PairCapacityExceeded:
            if attempt == retries:
                raise
    raise AssertionError("unreachable")


def dp_train_step(params: Sequence[torch.Tensor], optimizer, cameras: Sequence,
                  loss_fn: Callable[[object], torch.Tensor], bucket: GradBucket) -> torch.Tensor:
    """One data-parallel step over a camera batch.

    cameras: the GLOBAL batch of this step (same on every rank); rank r handles cameras r, r+world, ...
    loss_fn(cam) -> scalar loss of one camera (render through GaussianRasterizer + photometric loss).
    Per-camera losses are averaged over the global batch; returns this rank's share of that mean."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    for p in params:
        p.grad = None
    total = None
    for cam in cameras[rank::world]:
        loss = backward_or_retry(lambda c: loss_fn(c) / len(cameras), cam)
        total = loss.detach() if total is None else total + loss.detach()
    bucket.pack()
    bucket.all_reduce(average=False)         # losses are already divided by the global batch size
    bucket.unpack()
    optimizer.step()
    return total if total is not None else torch.zeros((), device=bucket.flat.device)
