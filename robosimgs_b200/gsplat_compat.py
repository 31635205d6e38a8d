"""gsplat-style ``rasterization()`` shim over the same kernels.

The reference delegates background reconstruction to Nerfstudio (/root/reference/README.md:75),
whose splatfacto model calls ``gsplat.rasterization``; this adapter maps that signature onto the
inria-style operator this repo implements:

    viewmats[C,4,4] (world->camera, OpenCV axes) + Ks[C,3,3]  ->  transposed view / full-projection
    matrices, tan(fov/2), camera centre;  quaternions normalised here (gsplat does it inside);
    near_plane is passed down (default 0.01 instead of the inria 0.2);  alphas = 1 - final T.

Only what the kernels implement is accepted -- everything else raises NotImplementedError rather than
silently rendering something different: render_mode "RGB", rasterize_mode "classic", eps2d 0.3,
tile_size 16, pinhole cameras, no far-plane / radius clipping, 3-channel colours.  ``meta["means2d"]`` is the
[C, N, 2] screen-space gradient carrier (pixel units, like gsplat's) that splatfacto's densification strategy
reads; ``alphas`` carry NO gradient (a loss term through the alpha image -- e.g. a learnt background blend --
gets none): the kernels' adjoint covers the colour image only.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .rasterizer import GaussianRasterizationSettings, _RasterizeGaussians


def _camera_matrices(viewmat: torch.Tensor, K: torch.Tensor, K_host, width: int, height: int, znear: float,
                     zfar: float):
    """K_host: the same intrinsics as Python floats (fx, fy) -- tan(fov/2) is a host scalar of the C ABI, and reading
    it from a CUDA `K` here would cost one device->host sync per camera (rasterization() fetches all cameras' Ks
    with a single copy instead)."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    tanfovx = float(width / (2.0 * K_host[0]))
    tanfovy = float(height / (2.0 * K_host[1]))
    P = torch.zeros(4, 4, dtype=viewmat.dtype, device=viewmat.device)
    P[0, 0] = 2.0 * fx / width
    P[1, 1] = 2.0 * fy / height
    # gsplat pixel k has its centre at k + 0.5, the rasterizer's at k: u - 0.5 = fx x/z + cx - 0.5
    P[0, 2] = (2.0 * cx - width) / width
    P[1, 2] = (2.0 * cy - height) / height
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    full = P @ viewmat
    campos = -viewmat[:3, :3].T @ viewmat[:3, 3]
    return viewmat.T.contiguous(), full.T.contiguous(), campos.contiguous(), tanfovx, tanfovy


def rasterization(means: torch.Tensor, quats: torch.Tensor, scales: torch.Tensor, opacities: torch.Tensor,
                  colors: torch.Tensor, viewmats: torch.Tensor, Ks: torch.Tensor, width: int, height: int,
                  near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0,
                  eps2d: float = 0.3, sh_degree: Optional[int] = None, packed: bool = True, tile_size: int = 16,
                  backgrounds: Optional[torch.Tensor] = None, render_mode: str = "RGB", sparse_grad: bool = False,
                  absgrad: bool = False, rasterize_mode: str = "classic", channel_chunk: int = 32,
                  distributed: bool = False, camera_model: str = "pinhole", **unsupported
                  ) -> Tuple[torch.Tensor, torch.Tensor, Dict]:
    """Returns (render_colors [C,H,W,3], render_alphas [C,H,W,1], meta)."""
    for name, ok in (("render_mode", render_mode == "RGB"), ("rasterize_mode", rasterize_mode == "classic"),
                     ("eps2d", abs(eps2d - 0.3) < 1e-12), ("tile_size", tile_size == 16),
                     ("camera_model", camera_model == "pinhole"), ("radius_clip", radius_clip == 0.0),
                     ("far_plane", far_plane >= 1e9), ("distributed", not distributed),
                     ("sparse_grad", not sparse_grad), ("absgrad", not absgrad)):
        if not ok:
            raise NotImplementedError(f"rasterization(): {name} value not supported by the b200gs kernels")
    if unsupported:
        raise NotImplementedError(f"rasterization(): unsupported arguments {sorted(unsupported)}")
    if colors.shape[-1] != 3:
        raise NotImplementedError("rasterization(): only 3-channel colours are supported")
    N = means.shape[0]
    C_ = viewmats.shape[0]
    dev = means.device
    quats_n = quats / quats.norm(dim=-1, keepdim=True)
    opac = opacities.reshape(N, 1)
    if sh_degree is None:
        shs, cols = means.new_empty(0), colors.reshape(N, 3)
    else:
        shs, cols = colors.reshape(N, -1, 3), means.new_empty(0)
    empty = means.new_empty(0)
    outs, alphas, radii_all = [], [], []
    K_host = Ks.detach()[:, [0, 1], [0, 1]].double().cpu().tolist()       # one copy for all cameras
    # meta["means2d"] [C, N, 2]: gradient carrier in PIXEL units, as gsplat's densification strategies read it
    # (info["means2d"].retain_grad() ... .grad); the kernels report the screen-space gradient NDC-scaled
    # (x 0.5 W, x 0.5 H -- the public inria convention), so the carrier enters scaled by (2/W, 2/H).
    means2d = torch.zeros((C_, N, 2), dtype=means.dtype, device=dev, requires_grad=torch.is_grad_enabled())
    to_ndc = torch.tensor([2.0 / width, 2.0 / height, 0.0], dtype=means.dtype, device=dev)
    for c in range(C_):
        view_t, proj_t, campos, tfx, tfy = _camera_matrices(viewmats[c].float(), Ks[c].float(), K_host[c], width, height,
                                                            max(near_plane, 1e-4), 1000.0)
        bg = backgrounds[c].float() if backgrounds is not None else torch.zeros(3, device=dev)
        rs = GaussianRasterizationSettings(int(height), int(width), tfx, tfy, bg, 1.0, view_t, proj_t,
                                           0 if sh_degree is None else int(sh_degree), campos, False, False)
        carrier = torch.nn.functional.pad(means2d[c], (0, 1)) * to_ndc
        color, radii, alpha = _RasterizeGaussians.apply(means, carrier, shs, cols, opac, scales, quats_n, empty, rs,
                                                        torch.is_grad_enabled(), float(near_plane), True)
        outs.append(color.permute(1, 2, 0))
        alphas.append(alpha[..., None])
        radii_all.append(radii)
    meta = {"radii": torch.stack(radii_all), "means2d": means2d, "width": width, "height": height, "tile_size": 16,
            "n_cameras": C_}
    return torch.stack(outs), torch.stack(alphas), meta
