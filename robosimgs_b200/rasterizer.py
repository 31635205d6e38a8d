"""PyTorch operator surface of the B200 rasterizer -- a drop-in for the Python package of the public
diff-gaussian-rasterization that BASELINE.json:north_star names (``GaussianRasterizationSettings``,
``GaussianRasterizer.forward/markVisible``, ``rasterize_gaussians``).  The reference
repo itself only delegates 3DGS reconstruction/rendering (/root/reference/README.md:75, :29, :85),
so names, argument meaning and error behaviour mirror that public interface.

Everything below is host-side marshalling; all arithmetic happens in libb200gs.so through the C ABI
in include/b200gs.h.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _cabi


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    antialiasing: bool = False


RasterizationSettings = GaussianRasterizationSettings


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _opt(t: Optional[torch.Tensor]):
    """The public interface passes unused optionals as empty tensors; normalise to None."""
    if t is None or t.numel() == 0:
        return None
    return t


def _f32c(t: Optional[torch.Tensor], name: str, device) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if t.device != device:
        raise ValueError(f"{name} is on {t.device}, expected {device}")
    return t.contiguous()


class _ScratchPool:
    """Per (device, stream) free-lists of byte tensors for the library's scratch buffers.

    The scratch sizes vary from frame to frame (they follow the pair count D); routing them through
    torch's caching allocator makes it split/merge large blocks for dozens of iterations (observed:
    60 cudaMallocs and 14 GB reserved before it settles).  Buffers are instead leased for the life
    of one forward call -- or of its autograd node when gradients are needed -- and handed back
    afterwards; reuse on the same stream is ordered by the stream itself."""

    def __init__(self):
        self.free: dict = {}

    @staticmethod
    def _round(n: int) -> int:
        g = 1 << max(int(n).bit_length() - 3, 12)      # keep 3 significant bits of head-room
        return (int(n) + g - 1) // g * g

    def acquire(self, key, nbytes: int, device) -> torch.Tensor:
        lst = self.free.setdefault(key, [])
        best = None
        for i, t in enumerate(lst):
            if t.numel() >= nbytes and (best is None or t.numel() < lst[best].numel()):
                best = i
        if best is not None:
            return lst.pop(best)
        if lst:                                          # grow: drop the largest too-small buffer
            lst.pop(max(range(len(lst)), key=lambda i: lst[i].numel()))
        return torch.empty(self._round(nbytes), dtype=torch.uint8, device=device)

    def release(self, key, t: torch.Tensor) -> None:
        self.free.setdefault(key, []).append(t)

    def clear(self) -> None:
        self.free.clear()


_POOL = _ScratchPool()


class _Lease:
    """Scratch buffers of one forward (or backward) call, handed to the library as B200GSAlloc
    resize callbacks (the public binding's resizeFunctional).  The callbacks close over a plain
    dict, not over the lease, so there is no reference cycle: dropping the last reference (e.g. the
    autograd node dying) returns the buffers to the pool immediately."""

    def __init__(self, device, kinds, tag=None):
        self.device = device
        self.stream = torch.cuda.current_stream(device).cuda_stream
        self.tensors = tensors = {}
        # tag (GaussianRasterizer.scratch_tag, set by sweep.SceneRenderer per frame slot) gives a caller private
        # scratch: buffers used inside a captured CUDA graph must never be handed to eager calls on the same stream
        keys = {k: (device.index, self.stream, k, tag) for k in kinds}
        self._keys = keys

        def make(kind):
            def resize(_ctx, nbytes):
                t = tensors.get(kind)
                if t is None or t.numel() < nbytes:
                    if t is not None:
                        _POOL.release(keys[kind], t)
                    t = tensors[kind] = _POOL.acquire(keys[kind], int(nbytes), device)
                return t.data_ptr()
            return _cabi.RESIZE_FN(resize)

        self._cbs = {k: make(k) for k in kinds}
        self.allocs = {k: _cabi.B200GSAlloc(None, cb) for k, cb in self._cbs.items()}

    def tensor(self, kind) -> torch.Tensor:
        t = self.tensors.get(kind)
        return t if t is not None else torch.empty(0, dtype=torch.uint8, device=self.device)

    def release(self) -> None:
        for kind, t in self.tensors.items():
            _POOL.release(self._keys[kind], t)
        self.tensors.clear()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


# Pair-capacity hints (see B200GSParams.pair_capacity_hint): a slowly decaying maximum of the pair
# counts of recent frames with the same (device, P, H, W), plus head-room -- so both a smooth camera
# path and random training views stay under the hint.  Purely a performance hint: the library redoes
# the binning stage exactly if a frame needs more.  The two rules themselves (head-room, bin size from the
# splat extent) live in the library -- b200gs_policy_pair_capacity / b200gs_policy_bin_shift, the same functions a
# B200GSContext applies for C callers (include/b200gs.h) -- this layer only keeps the per-key state next to autograd.
_PAIR_HINTS: dict = {}
SPECULATE_PAIR_CAPACITY = True


def _params(P, M, rs: GaussianRasterizationSettings, hint: int = 0, near_plane: float = 0.0,
            flags: int = 0) -> _cabi.B200GSParams:
    return _cabi.B200GSParams(int(P), int(rs.sh_degree), int(M), int(rs.image_height), int(rs.image_width),
                              float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier),
                              int(bool(rs.prefiltered)), int(bool(rs.debug)), float(near_plane), int(flags), int(hint))


# Bin-size policy (B200GSParams.flags bits 8..11).  The library's automatic bin size only looks at the image
# (largest bins that keep the sort key narrow: 128 px at 1080p) -- right for scenes of large splats, but with
# small splats a tile then walks thousands of records of its bin to find the few that touch it.  On the first
# frame of a (device, P, H, W) the pairs-per-touching-Gaussian ratio D / #(radii > 0) ~ (1 + extent/bin)^2 gives
# the typical splat extent on screen; later frames ask for bins of about three extents (32..256 px).  Results
# do not depend on the bin size.  One reduction + sync on that first frame, re-checked every 256 synchronous calls.
_BIN_POLICY: dict = {}
ADAPT_BIN_SIZE = True


def _default_bin_shift(H: int, W: int) -> int:
    gx, gy = (W + 15) // 16, (H + 15) // 16
    s = 0
    while s < 3 and ((gx + (1 << s) - 1) >> s) * ((gy + (1 << s) - 1) >> s) > 255:
        s += 1
    return s


def _bin_flags(hint_key, H, W):
    pol = _BIN_POLICY.setdefault(hint_key, {"shift": -1, "calls": 0})
    used = pol["shift"] if pol["shift"] >= 0 else _default_bin_shift(H, W)
    return pol, used, ((pol["shift"] + 1) << 8 if pol["shift"] >= 0 else 0)


def _adapt_bin_size(pol, used, hint_key, D, radii, coverage=None):
    """coverage: callable returning the mean of (1 - final transmittance) over the frame, or None."""
    pol["calls"] += 1
    if not ADAPT_BIN_SIZE or D <= 0 or not (pol["calls"] == 1 or pol["calls"] % 256 == 0):
        return
    touch = int((radii > 0).sum())
    if touch <= 0:
        return
    # bins of about three splat extents (32..256 px); one size coarser when 128 px came out and the frame saturates
    # everywhere (C3: 0.372 -> 0.356 ms) -- the rule is b200gs_policy_bin_shift; the coverage (a device reduction)
    # is only evaluated when it can matter
    L = _cabi.lib()
    new = int(L.b200gs_policy_bin_shift(int(D), touch, int(used), -1.0))
    if new == 3 and coverage is not None:
        new = int(L.b200gs_policy_bin_shift(int(D), touch, int(used), float(coverage())))
    if new != used:
        pol["shift"] = new
        _PAIR_HINTS.pop(hint_key, None)          # the pair count changes with the bin size


class PairTicket:
    """Receipt of a forward call made with the pair-count check DEFERRED (forward-only sweeps).

    The library launched the frame for `hint` pair slots and returned without waiting for the pair
    count D; D arrives in a pinned word once the frame's stream has passed the call.  ``ok()`` waits
    for that point and tells whether the frame is complete (D <= hint).  A frame that is not must be
    rendered again (``GaussianRasterizer.forward`` does it exactly) -- with the decaying-maximum hints
    this happens only on abrupt view changes."""
    _free_words: list = []

    def __init__(self, hint: int, hint_key, word=None):
        self.hint, self.hint_key = int(hint), hint_key
        self._own_word = word is None
        if word is None:
            word = PairTicket._free_words.pop() if PairTicket._free_words else torch.zeros(1, dtype=torch.int32).pin_memory()
        self.word = word
        self.event = None
        self._ok = None
        self.pairs = None

    def ok(self) -> bool:
        if self._ok is None:
            if self.event is not None:
                self.event.synchronize()
            D = self.pairs = int(self.word[0])
            self._ok = self.hint <= 0 or D <= self.hint
            last = _PAIR_HINTS.get(self.hint_key, 0)
            _PAIR_HINTS[self.hint_key] = max(D, int(last * 0.97))
            if self._own_word:
                PairTicket._free_words.append(self.word)
            self.word = None
        return self._ok


class DeferOptions(list):
    """Ticket box of a deferred forward call (the ticket is appended) with optional overrides:
    capacity -- explicit pair capacity instead of the decaying-maximum hint (a captured CUDA graph
    needs launch arguments that do not change from frame to frame); word -- caller-owned pinned int32
    word that receives the pair count; record_event -- False inside stream capture (the caller records
    its own event after replaying the graph)."""

    def __init__(self, capacity: int = 0, word=None, record_event: bool = True, train: bool = False, rgb8=None):
        super().__init__()
        self.capacity, self.word, self.record_event = int(capacity), word, bool(record_event)
        self.train = bool(train)      # the frame keeps its adjoint state; backward validates the ticket
        # rgb8: CUDA uint8 tensor [H, W, 3] -- the compositing kernel writes the finished 8-bit frame straight into it
        # (B200GS_OUT_RGB8; bit-identical to export_rgb8 of the fp32 frame) and the call returns it as `color`
        self.rgb8 = rgb8


class PairCapacityExceeded(_cabi.B200GSError):
    """A training step made with ``GaussianRasterizer.defer_pair_check`` produced more pairs than the speculative
    capacity: the frame the loss saw was incomplete and so are the gradients.  The capacity hint has been raised;
    run the step again (robosimgs_b200.train.backward_or_retry does)."""


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings, grad_mode=True, near_plane=0.0, want_alpha=False, ticket_box=None,
                scratch_tag=None):
        L = _cabi.lib()
        rs = raster_settings
        if getattr(rs, "antialiasing", False):
            raise NotImplementedError("antialiasing=True is not implemented by this rasterizer")
        dev = means3D.device
        if dev.type != "cuda":
            raise _cabi.B200GSError("b200gs rasterizer needs CUDA tensors; there is no CPU fallback")
        means3D = _f32c(means3D, "means3D", dev)
        sh = _f32c(_opt(sh), "shs", dev)
        colors_precomp = _f32c(_opt(colors_precomp), "colors_precomp", dev)
        opacities = _f32c(opacities, "opacities", dev)
        scales = _f32c(_opt(scales), "scales", dev)
        rotations = _f32c(_opt(rotations), "rotations", dev)
        cov3Ds_precomp = _f32c(_opt(cov3Ds_precomp), "cov3D_precomp", dev)
        bg = _f32c(rs.bg, "bg", dev)
        view = _f32c(rs.viewmatrix, "viewmatrix", dev)
        proj = _f32c(rs.projmatrix, "projmatrix", dev)
        campos = _f32c(rs.campos, "campos", dev)

        P = means3D.shape[0]
        M = 0 if sh is None else (sh.shape[1] if sh.dim() == 3 else sh.reshape(P, -1, 3).shape[1])
        H, W = int(rs.image_height), int(rs.image_width)
        hint_key = (dev.index, P, H, W)
        last_D = _PAIR_HINTS.get(hint_key, 0) if SPECULATE_PAIR_CAPACITY and not rs.debug else 0
        hint = int(L.b200gs_policy_pair_capacity(int(last_D)))
        # deferred pair check (forward-only callers that pass a ticket box): only once a hint exists
        ticket = None
        if ticket_box is not None:
            opts = ticket_box if isinstance(ticket_box, DeferOptions) else DeferOptions()
            if torch.is_grad_enabled() and any(ctx.needs_input_grad) and not opts.train:
                raise _cabi.B200GSError("the deferred pair check is for forward-only rendering (no_grad)")
            if opts.capacity > 0:
                hint = opts.capacity
            ticket = PairTicket(hint if P > 0 else 0, hint_key, opts.word)
            ticket_box.append(ticket)
        defer = ticket is not None and hint > 0 and P > 0
        pol, shift_used, bin_flags = _bin_flags(hint_key, H, W)
        # a deferred frame is forward-only by construction: the per-pixel state of the adjoint (20 B per pixel) is not
        # written unless the caller wants the alpha channel, which is read from it
        # -- and so is a blocking frame nobody can differentiate (no_grad, or no input requires a gradient), unless
        # the bin-size policy is about to read the frame's coverage from that state (first and every 256th call)
        needs_grad = bool(grad_mode) and any(ctx.needs_input_grad)
        policy_looks = ADAPT_BIN_SIZE and not rs.debug and ((pol["calls"] + 1) == 1 or (pol["calls"] + 1) % 256 == 0)
        if defer:
            keep_state = want_alpha or opts.train
        else:
            keep_state = want_alpha or needs_grad or policy_looks
        fwd_flags = (_cabi.DEFER_PAIR_CHECK if defer else 0) | (0 if keep_state else _cabi.FORWARD_ONLY)
        rgb8 = opts.rgb8 if ticket_box is not None else None
        if rgb8 is not None:
            if rgb8.dtype != torch.uint8 or tuple(rgb8.shape) != (H, W, 3) or rgb8.device != dev or not rgb8.is_contiguous():
                raise _cabi.B200GSError("rgb8 output must be a contiguous CUDA uint8 tensor of shape [H, W, 3]")
            if opts.train:
                raise _cabi.B200GSError("an 8-bit frame carries no gradient: rgb8 output is for forward-only rendering")
            fwd_flags |= _cabi.OUT_RGB8
        prm = _params(P, M, rs, hint, near_plane, fwd_flags | bin_flags)
        color = rgb8 if rgb8 is not None else torch.empty((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        num_rendered = C.c_int32(0)
        nr_ptr = C.c_void_p(ticket.word.data_ptr()) if defer else C.cast(C.pointer(num_rendered), C.c_void_p)
        with torch.cuda.device(dev):
            lease = _Lease(dev, ("geom", "binning", "img"), scratch_tag)
            stream = C.c_void_p(lease.stream)
            _cabi.check(L.b200gs_forward(
                C.byref(prm), _ptr(bg), _ptr(view), _ptr(proj), _ptr(campos), _ptr(means3D), _ptr(sh),
                _ptr(colors_precomp), _ptr(opacities), _ptr(scales), _ptr(rotations),
                _ptr(cov3Ds_precomp), _ptr(color), _ptr(radii), lease.allocs["geom"], lease.allocs["binning"],
                lease.allocs["img"], nr_ptr, stream))
            if defer:
                if opts.record_event:
                    ticket.event = torch.cuda.Event()
                    ticket.event.record(torch.cuda.current_stream(dev))
            elif ticket is not None:
                ticket.word[0] = int(num_rendered.value)     # synchronous path: already exact
                ticket.hint = 0
            alpha = None
            if want_alpha:
                alpha = torch.empty((H, W), dtype=torch.float32, device=dev)
                _cabi.check(L.b200gs_extract_alpha(_ptr(lease.tensor("img")), C.c_int32(H), C.c_int32(W),
                                                   _ptr(alpha), stream))
        ctx.raster_settings = rs
        ctx.near_plane = near_plane
        # (a deferred training frame: the adjoint only needs to know that there are pairs; it validates the ticket)
        ctx.num_rendered = int(hint) if defer else int(num_rendered.value)
        ctx.ticket = ticket if (defer and opts.train) else None
        ctx.bin_flags = bin_flags
        if not defer:
            _PAIR_HINTS[hint_key] = max(ctx.num_rendered, int(last_D * 0.97))
            if not rs.debug:
                def coverage():
                    a = torch.empty((H, W), dtype=torch.float32, device=dev)
                    with torch.cuda.device(dev):
                        _cabi.check(L.b200gs_extract_alpha(_ptr(lease.tensor("img")), C.c_int32(H), C.c_int32(W),
                                                           _ptr(a), stream))
                    return float(a.mean())
                _adapt_bin_size(pol, shift_used, hint_key, ctx.num_rendered, radii, coverage)
        ctx.M = M
        ctx.in_shapes = (None if sh is None else tuple(sh.shape), tuple(opacities.shape))
        ctx.present = (sh is not None, colors_precomp is not None, scales is not None,
                       cov3Ds_precomp is not None)
        if grad_mode and any(ctx.needs_input_grad):
            z = means3D.new_empty(0)
            ctx.lease = lease      # scratch stays leased until the autograd node dies
            ctx.save_for_backward(means3D, z if sh is None else sh, z if colors_precomp is None else colors_precomp,
                                  opacities, z if scales is None else scales, z if rotations is None else rotations,
                                  z if cov3Ds_precomp is None else cov3Ds_precomp, radii, lease.tensor("geom"),
                                  lease.tensor("binning"), lease.tensor("img"), bg, view, proj, campos)
        else:
            lease.release()        # forward-only: the next call on this stream may reuse the scratch
        ctx.mark_non_differentiable(radii)
        if want_alpha:
            ctx.mark_non_differentiable(alpha)
            return color, radii, alpha
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, *_unused):
        L = _cabi.lib()
        rs = ctx.raster_settings
        (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, radii, geom,
         binning, img, bg, view, proj, campos) = ctx.saved_tensors
        has_sh, has_col, has_sr, has_cov = ctx.present
        sh = sh if has_sh else None
        colors_precomp = colors_precomp if has_col else None
        scales = scales if has_sr else None
        rotations = rotations if has_sr else None
        cov3Ds_precomp = cov3Ds_precomp if has_cov else None
        dev = means3D.device
        P = means3D.shape[0]
        grad_out_color = grad_out_color.to(torch.float32).contiguous()
        e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        g_means3D, g_means2D, g_opac = e(P, 3), e(P, 3), e(P, 1)
        g_sh = e(P, ctx.M, 3) if has_sh else None
        g_col = e(P, 3) if has_col else None
        g_scales = e(P, 3) if has_sr else None
        g_rots = e(P, 4) if has_sr else None
        g_cov = e(P, 6) if has_cov else None
        prm = _params(P, ctx.M, rs, 0, ctx.near_plane, ctx.bin_flags)
        with torch.cuda.device(dev):
            lease = _Lease(dev, ("scratch",))
            stream = C.c_void_p(lease.stream)
            _cabi.check(L.b200gs_backward(
                C.byref(prm), _ptr(bg), _ptr(view), _ptr(proj), _ptr(campos), _ptr(means3D), _ptr(sh),
                _ptr(colors_precomp), _ptr(opacities), _ptr(scales), _ptr(rotations),
                _ptr(cov3Ds_precomp), _ptr(radii), _ptr(geom), _ptr(binning), _ptr(img),
                C.c_int32(ctx.num_rendered), _ptr(grad_out_color), _ptr(g_means3D), _ptr(g_means2D),
                _ptr(g_sh), _ptr(g_col), _ptr(g_opac), _ptr(g_scales), _ptr(g_rots), _ptr(g_cov),
                lease.allocs["scratch"], stream))
            lease.release()
        if ctx.ticket is not None and not ctx.ticket.ok():
            # everything above was launched for an incomplete frame (harmless: no kernel indexes past its buffers);
            # the host only learns it here, with the whole step already queued -- it never waited inside the step
            raise PairCapacityExceeded(f"{ctx.ticket.pairs} pairs, capacity {ctx.ticket.hint}: run the step again")
        # order of the public interface: means3D, means2D, sh, colors_precomp, opacities, scales,
        # rotations, cov3Ds_precomp, raster_settings
        sh_shape, opac_shape = ctx.in_shapes        # forward accepts shs [P, M*3] and opacities [P]: mirror them
        if g_sh is not None and sh_shape is not None:
            g_sh = g_sh.view(sh_shape)
        g_opac = g_opac.view(opac_shape)
        return g_means3D, g_means2D, g_sh, g_col, g_opac, g_scales, g_rots, g_cov, None, None, None, None, None, None


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings, scratch_tag=None, ticket_box=None):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, torch.is_grad_enabled(), 0.0, False, ticket_box,
                                     scratch_tag)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings
        self.scratch_tag = None      # not None: this rasterizer leases PRIVATE scratch (see _Lease)
        # Training without a host wait inside the step (opt-in; the public operator's contract is a complete frame on
        # return, so the default is off): forward launches the frame for the speculative pair capacity and returns at
        # once, backward launches the adjoint and only THEN checks the pair count that has long arrived in a pinned
        # word.  If it exceeded the capacity (an abrupt view change) backward raises PairCapacityExceeded and the
        # caller repeats the step; `last_ticket` lets a caller that never runs backward validate the frame.
        self.defer_pair_check = False
        self.last_ticket = None

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """Boolean mask of the Gaussians that pass the near-plane cull (view-space z > 0.2)."""
        L = _cabi.lib()
        rs = self.raster_settings
        with torch.no_grad():
            dev = positions.device
            if dev.type != "cuda":
                raise _cabi.B200GSError("b200gs rasterizer needs CUDA tensors; there is no CPU fallback")
            pos = _f32c(positions, "positions", dev)
            view = _f32c(rs.viewmatrix, "viewmatrix", dev)
            proj = _f32c(rs.projmatrix, "projmatrix", dev)
            out = torch.empty((pos.shape[0],), dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                _cabi.check(L.b200gs_mark_visible(C.c_int32(pos.shape[0]), _ptr(pos), _ptr(view), _ptr(proj),
                                                  _ptr(out), stream))
        return out.bool()

    def forward_deferred(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                         rotations=None, cov3D_precomp=None, options: Optional[DeferOptions] = None):
        """Forward-only render that never blocks the host: returns ``(color, radii, ticket)``.  The frame
        may be consumed once ``ticket.ok()`` is True; if it is False (pair count above the speculative
        capacity) call ``forward`` again for this frame.  Use under ``torch.no_grad()``."""
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        e = means3D.new_empty(0)
        box = options if options is not None else DeferOptions()
        # forward-only by construction: nothing is saved for a backward pass and the frame may be incomplete
        # until the ticket is validated, so the result must never carry a grad_fn
        det = lambda t: e if t is None else t.detach()
        with torch.no_grad():
            color, radii = _RasterizeGaussians.apply(
                means3D.detach(), means2D.detach(), det(shs), det(colors_precomp), opacities.detach(), det(scales),
                det(rotations), det(cov3D_precomp), self.raster_settings, False, 0.0, False, box, self.scratch_tag)
        return color, radii, box[-1]

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = means3D.new_empty(0)
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        box = None
        if self.defer_pair_check and torch.is_grad_enabled():
            box = DeferOptions(train=True)
        out = rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                  cov3D_precomp, rs, self.scratch_tag, box)
        if box is not None:
            # a frame whose backward never ran (evaluation with grad enabled) still gives its pinned word back and
            # feeds the capacity hint, as soon as its stream has passed it
            old = self.last_ticket
            if old is not None and old._ok is None and (old.event is None or old.event.query()):
                old.ok()
            self.last_ticket = box[-1]
        return out


def export_rgb8(color: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """color[3,H,W] fp32 (CUDA) -> uint8 [H,W,3] RGB on the same device and stream (clamp, *255,
    round) -- the frame format a datagen sweep stores; a 1080p frame is 6.2 MB instead of 24.9 MB."""
    L = _cabi.lib()
    if color.device.type != "cuda":
        raise _cabi.B200GSError("b200gs rasterizer needs CUDA tensors; there is no CPU fallback")
    color = _f32c(color.detach(), "color", color.device)
    _, H, W = color.shape
    if out is None:
        out = torch.empty((H, W, 3), dtype=torch.uint8, device=color.device)
    with torch.cuda.device(color.device):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _cabi.check(L.b200gs_export_rgb8(_ptr(color), C.c_int32(H), C.c_int32(W), _ptr(out), stream))
    return out
