"""Seeded synthetic scenes for the BASELINE.json configs (SURVEY.md 8(d) "Synthetic inputs").

Everything is generated on the CPU with ``torch.Generator().manual_seed(seed)`` so the CPU
oracle and the GPU path see bit-identical inputs.  Tensors are *post-activation* (opacity after
sigmoid, scale after exp, unit quaternions wxyz) -- the layout GaussianRasterizer.forward takes.
"""
from __future__ import annotations

import math
from typing import NamedTuple

import torch

from .cameras import Camera, camera_look_at, orbit_cameras, reference_six_views

SH_C0 = 0.28209479177387814


class Scene(NamedTuple):
    means3D: torch.Tensor    # [P,3]
    shs: torch.Tensor        # [P,M,3]
    opacities: torch.Tensor  # [P,1]
    scales: torch.Tensor     # [P,3]
    rotations: torch.Tensor  # [P,4] wxyz, unit
    sh_degree: int

    @property
    def P(self) -> int:
        return self.means3D.shape[0]

    def to(self, device) -> "Scene":
        return Scene(*(t.to(device) if torch.is_tensor(t) else t for t in self))


def _rand_quats(n, g):
    q = torch.randn(n, 4, generator=g)
    return q / q.norm(dim=1, keepdim=True)


def _sh_from_rgb(rgb, degree, g, rest_sigma=0.1):
    P = rgb.shape[0]
    M = (degree + 1) ** 2
    shs = torch.zeros(P, M, 3)
    shs[:, 0] = (rgb - 0.5) / SH_C0
    if M > 1:
        shs[:, 1:] = torch.randn(P, M - 1, 3, generator=g) * rest_sigma
    return shs


def cube_scene(P: int = 10_000, seed: int = 1, degree: int = 0):
    """C1: 10k-Gaussian cube, 256x256, fov 50 deg, eye (0,0,4) -> origin, bg 0."""
    g = torch.Generator().manual_seed(seed)
    means = torch.rand(P, 3, generator=g) * 2 - 1
    lo, hi = math.log(0.01), math.log(0.05)
    scales = torch.exp(torch.rand(P, 3, generator=g) * (hi - lo) + lo)
    rots = _rand_quats(P, g)
    opac = torch.sigmoid(torch.randn(P, 1, generator=g))
    rgb = torch.rand(P, 3, generator=g)
    sc = Scene(means, _sh_from_rgb(rgb, degree, g), opac, scales, rots, degree)
    cam = camera_look_at((0, 0, 4), (0, 0, 0), (0, 1, 0), 50.0, 256, 256)
    return sc, cam


def tabletop_scene(P: int = 200_000, seed: int = 2, degree: int = 3, resolution: int = 800):
    """C2: 70% flat splats on a 2x2 m slab, 30% in 5 object blobs; the reference's six views."""
    g = torch.Generator().manual_seed(seed)
    n_plane = int(P * 0.7)
    n_obj = P - n_plane
    pm = torch.empty(n_plane, 3)
    pm[:, :2] = torch.rand(n_plane, 2, generator=g) * 2 - 1
    pm[:, 2] = (torch.rand(n_plane, generator=g) * 2 - 1) * 0.02
    ps = torch.empty(n_plane, 3)
    ps[:, :2] = torch.rand(n_plane, 2, generator=g) * 0.02 + 0.01
    ps[:, 2] = torch.rand(n_plane, generator=g) * 0.002 + 0.001
    pq = torch.zeros(n_plane, 4)
    ang = torch.rand(n_plane, generator=g) * math.pi
    pq[:, 0], pq[:, 3] = torch.cos(ang / 2), torch.sin(ang / 2)   # rotation about z
    centers = torch.rand(5, 3, generator=g) * torch.tensor([1.4, 1.4, 0.0]) + torch.tensor([-0.7, -0.7, 0.15])
    which = torch.randint(0, 5, (n_obj,), generator=g)
    om = centers[which] + torch.randn(n_obj, 3, generator=g) * 0.1
    lo, hi = math.log(0.005), math.log(0.03)
    osc = torch.exp(torch.rand(n_obj, 3, generator=g) * (hi - lo) + lo)
    oq = _rand_quats(n_obj, g)
    means = torch.cat([pm, om]); scales = torch.cat([ps, osc]); rots = torch.cat([pq, oq])
    opac = torch.sigmoid(torch.randn(P, 1, generator=g) + 1.0)
    rgb = torch.rand(P, 3, generator=g)
    sc = Scene(means, _sh_from_rgb(rgb, degree, g), opac, scales, rots, degree)
    cams = reference_six_views((0.0, 0.0, 0.1), 1.0, resolution)
    return sc, cams


def room_scene(P: int = 1_000_000, seed: int = 3, degree: int = 3, width: int = 1920,
               height: int = 1080):
    """C3: shell of a 6x5x3 m box (80%) + clutter blobs (20%); camera inside, fov_x 70 deg."""
    g = torch.Generator().manual_seed(seed)
    n_shell = int(P * 0.8)
    n_cl = P - n_shell
    half = torch.tensor([3.0, 2.5, 1.5])
    u = torch.rand(n_shell, 3, generator=g) * 2 - 1
    face = torch.randint(0, 6, (n_shell,), generator=g)
    axis, sign = face // 2, (face % 2).float() * 2 - 1
    u[torch.arange(n_shell), axis] = sign
    shell = u * half + torch.randn(n_shell, 3, generator=g) * 0.01
    cc = (torch.rand(12, 3, generator=g) * 2 - 1) * half * 0.7
    which = torch.randint(0, 12, (n_cl,), generator=g)
    clutter = cc[which] + torch.randn(n_cl, 3, generator=g) * 0.25
    means = torch.cat([shell, clutter])
    ls = torch.randn(P, 3, generator=g) * 0.5 + math.log(0.02)
    scales = torch.exp(ls.clamp(math.log(0.003), math.log(0.2)))
    # flatten shell splats along the wall normal (surfels), like a trained scene
    flat = torch.ones(P, 3)
    flat[torch.arange(n_shell), axis] = 0.15
    scales = scales * flat
    rots = torch.zeros(P, 4)
    rots[:, 0] = 1.0
    rots[n_shell:] = _rand_quats(n_cl, g)
    jitter = _rand_quats(n_shell, g) * 0.05
    rots[:n_shell] = rots[:n_shell] + jitter
    rots = rots / rots.norm(dim=1, keepdim=True)
    opac = torch.sigmoid(torch.randn(P, 1, generator=g) * 1.5 + 1.0)
    rgb = torch.rand(P, 3, generator=g)
    sc = Scene(means, _sh_from_rgb(rgb, degree, g), opac, scales, rots, degree)
    cam = camera_look_at((-1.8, -1.2, 0.1), (2.0, 1.0, -0.2), (0, 0, 1), 70.0, width, height)
    return sc, cam


def room_target(width: int = 1920, height: int = 1080, seed: int = 33) -> torch.Tensor:
    """Fixed U[0,1] noise target for the train-step loss L = mean((img - target)^2)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(3, height, width, generator=g)


def sweep_scene(P: int = 3_000_000, n_cams: int = 64, seed: int = 4, width: int = 1920,
                height: int = 1080):
    """C4: the room generator scaled to P Gaussians + a seeded 64-camera inward orbit."""
    sc, _ = room_scene(P, seed, 3, width, height)
    cams = orbit_cameras(n_cams, (0.0, 0.0, 0.0), 2.0, 70.0, width, height, seed=44)
    return sc, cams


def mse_loss(img: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return ((img - target) ** 2).mean()


def settings_from_camera(cam: Camera, sh_degree: int, bg=(0.0, 0.0, 0.0), scale_modifier=1.0,
                         device="cpu", **kw):
    from .rasterizer import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx,
        tanfovy=cam.tanfovy, bg=torch.tensor(bg, dtype=torch.float32, device=device),
        scale_modifier=scale_modifier, viewmatrix=cam.viewmatrix.to(device),
        projmatrix=cam.projmatrix.to(device), sh_degree=sh_degree, campos=cam.campos.to(device),
        prefiltered=False, debug=False, **kw)
