"""Independent, differentiable PyTorch-CPU restatement of the 3DGS forward (SURVEY.md 8(c) oracle
spec), used ONLY to cross-check the hand-derived analytic backward of oracle/gs_oracle.c: gradients
here come from plain autograd in fp64.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/gs_oracle_impl.h).

Formulation: vectorised per-Gaussian projection; per tile a dense [pixels x Gaussians] alpha matrix
in (depth, index) order, transmittance by cumprod; the skip / stop rules of step 11 are applied as
(detached) masks so the composited value is exactly the sequential algorithm's.
"""
from __future__ import annotations

import math

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def sh_basis(deg, d):
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    b = [torch.full_like(x, C0)]
    if deg > 0:
        b += [-C1 * y, C1 * z, -C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
        if deg > 2:
            b += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
                  C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy),
                  C3[5] * z * (xx - yy), C3[6] * x * (xx - 3 * yy)]
    return torch.stack(b, dim=1)  # [P, nb]


def render(settings, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
           cov3D_precomp=None):
    """Returns (color[3,H,W], radii[P], means2D_ndc_proxy).  All inputs float64 CPU tensors.
    ``means2D_proxy`` is a zero tensor added to the pixel centre in NDC-scaled units so that its
    .grad equals the public implementation's dL_dmeans2D (pixel gradient * 0.5*W / 0.5*H)."""
    dt = means3D.dtype
    H, W = int(settings.image_height), int(settings.image_width)
    V = settings.viewmatrix.to(dt).T      # actual world->view (rows)
    Pm = settings.projmatrix.to(dt).T     # actual full projection
    campos = settings.campos.to(dt)
    bg = settings.bg.to(dt)
    tfx, tfy = float(settings.tanfovx), float(settings.tanfovy)
    mod = float(settings.scale_modifier)
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=dt)
    hom = torch.cat([means3D, ones], 1)
    t = hom @ V.T                              # [P,4] view space
    ph = hom @ Pm.T
    pw = 1.0 / (ph[:, 3] + 1e-7)
    ndc = ph[:, :2] * pw[:, None]
    vis = t[:, 2] > 0.2

    if cov3D_precomp is None:
        r, x, y, z = rotations[:, 0], rotations[:, 1], rotations[:, 2], rotations[:, 3]
        R = torch.stack([
            1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(P, 3, 3)
        Mx = R * (mod * scales)[:, None, :]
        Sigma = Mx @ Mx.transpose(1, 2)
    else:
        c = cov3D_precomp
        Sigma = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]],
                            dim=1).reshape(P, 3, 3)

    tz = torch.where(vis, t[:, 2], torch.ones_like(t[:, 2]))
    limx, limy = 1.3 * tfx, 1.3 * tfy
    txtz, tytz = t[:, 0] / tz, t[:, 1] / tz
    xin = ((txtz >= -limx) & (txtz <= limx)).to(dt)
    yin = ((tytz >= -limy) & (tytz <= limy)).to(dt)
    # clamped branch is treated as a constant w.r.t. autograd (public implementation's x/y_grad_mul)
    tx = t[:, 0] * xin + (txtz.clamp(-limx, limx) * tz).detach() * (1 - xin)
    ty = t[:, 1] * yin + (tytz.clamp(-limy, limy) * tz).detach() * (1 - yin)
    fx, fy = W / (2 * tfx), H / (2 * tfy)
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz), zero, fy / tz, -(fy * ty) / (tz * tz)], dim=1).reshape(P, 2, 3)
    Mjw = J @ V[:3, :3]
    cov2 = Mjw @ Sigma @ Mjw.transpose(1, 2)
    a = cov2[:, 0, 0] + 0.3
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    det_safe = torch.where(det == 0, torch.ones_like(det), det)
    A, B, Cc = c / det_safe, -b / det_safe, a / det_safe
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3 * torch.sqrt(lam)).detach()
    px = ((ndc[:, 0] + 1) * W - 1) * 0.5
    py = ((ndc[:, 1] + 1) * H - 1) * 0.5
    means2D_proxy = torch.zeros(P, 2, dtype=dt, requires_grad=True)
    px = px + means2D_proxy[:, 0] * (0.5 * W)
    py = py + means2D_proxy[:, 1] * (0.5 * H)

    gx, gy = (W + 15) // 16, (H + 15) // 16
    pxd, pyd = px.detach(), py.detach()
    x0 = torch.trunc((pxd - radius) / 16).clamp(0, gx).long()
    y0 = torch.trunc((pyd - radius) / 16).clamp(0, gy).long()
    x1 = torch.trunc((pxd + radius + 15) / 16).clamp(0, gx).long()
    y1 = torch.trunc((pyd + radius + 15) / 16).clamp(0, gy).long()
    ok = vis & (det != 0) & ((x1 - x0) * (y1 - y0) > 0)
    radii = torch.where(ok, radius, torch.zeros_like(radius)).to(torch.int32)

    if colors_precomp is None:
        d = means3D - campos[None]
        d = d / d.norm(dim=1, keepdim=True)
        deg = int(settings.sh_degree)
        nb = (deg + 1) ** 2
        rgb = (sh_basis(deg, d)[:, :, None] * shs[:, :nb, :]).sum(1) + 0.5
        rgb = torch.clamp(rgb, min=0.0)   # autograd gives zero grad where clamped
    else:
        rgb = colors_precomp

    depth = t[:, 2].detach()
    color = torch.zeros(3, H, W, dtype=dt)
    out_rows = []
    idx_ok = torch.nonzero(ok).flatten()
    for tyi in range(gy):
        row_tiles = []
        for txi in range(gx):
            sel = idx_ok[(x0[idx_ok] <= txi) & (x1[idx_ok] > txi) & (y0[idx_ok] <= tyi) & (y1[idx_ok] > tyi)]
            hx, hy = min(16, W - txi * 16), min(16, H - tyi * 16)
            ys, xs = torch.meshgrid(torch.arange(hy, dtype=dt) + tyi * 16, torch.arange(hx, dtype=dt) + txi * 16,
                                    indexing="ij")
            pxs, pys = xs.reshape(-1), ys.reshape(-1)
            npx = pxs.numel()
            if sel.numel() == 0:
                tile = bg[:, None].expand(3, npx)
                row_tiles.append(tile.reshape(3, hy, hx))
                continue
            # stable order: depth, then index
            order = torch.argsort(depth[sel], stable=True)
            sel = sel[order]
            dx = px[sel][None, :] - pxs[:, None]
            dy = py[sel][None, :] - pys[:, None]
            power = -0.5 * (A[sel][None] * dx * dx + Cc[sel][None] * dy * dy) - B[sel][None] * dx * dy
            alpha = opacities.reshape(-1)[sel][None] * torch.exp(power)
            # min(0.99, .) is NOT masked in the public adjoint: straight-through clamp
            alpha = alpha + (torch.clamp(alpha, max=0.99) - alpha).detach()
            keep = ((power <= 0) & (alpha >= 1.0 / 255)).detach()
            alpha = torch.where(keep, alpha, torch.zeros_like(alpha))
            one_m = 1 - alpha
            Tincl = torch.cumprod(one_m, dim=1)                       # T after each Gaussian
            Texcl = torch.cat([torch.ones(npx, 1, dtype=dt), Tincl[:, :-1]], dim=1)
            stop = (keep & (Tincl.detach() < 1e-4))
            stopped = (torch.cumsum(stop.to(torch.int32), dim=1) > 0)  # this one and all later ones
            wgt = torch.where(stopped, torch.zeros_like(alpha), alpha * Texcl)
            # final T: product over the contributing (non-stopped) Gaussians
            Tfin = torch.prod(torch.where(stopped, torch.ones_like(one_m), one_m), dim=1)
            col = wgt @ rgb[sel]                                        # [npx,3]
            tile = col.T + Tfin[None, :] * bg[:, None]
            row_tiles.append(tile.reshape(3, hy, hx))
        out_rows.append(torch.cat(row_tiles, dim=2))
    color = torch.cat(out_rows, dim=1)
    return color, radii, means2D_proxy
