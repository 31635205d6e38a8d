"""Camera convention adapter: OpenGL camera-to-world (+ intrinsics) -> rasterizer settings fields.

The reference repo states its camera wire format in two places, both OpenGL (camera looks down
-Z, +Y up) camera-to-world matrices with pinhole intrinsics:
  * Nerfstudio ``transforms.json`` / ``dataparser_transforms.json``
    (/root/reference/Articulation/utils/nerf2physic_utils.py:26-61; the GL->CV flip is at :16
    and :180),
  * the six synthetic views of the segmenter
    (/root/reference/Articulation/segmentation/interactive_segmenter.py:262-313; projection at
    :1436-1460, image row = H - v).
The rasterizer surface wants: transposed world->view matrix (OpenCV axes:
x right, y down, z forward), transposed full projection, tan(fov/2) and the camera centre.
"""
from __future__ import annotations

import json
import math
from typing import NamedTuple, Sequence

import numpy as np
import torch


class Camera(NamedTuple):
    """Fields named after GaussianRasterizationSettings; matrices are already transposed."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor   # [4,4] float32, transposed world->view
    projmatrix: torch.Tensor   # [4,4] float32, transposed full projection (proj @ view)
    campos: torch.Tensor       # [3] float32


def look_at_c2w_opengl(eye: Sequence[float], target: Sequence[float], up: Sequence[float]) -> np.ndarray:
    """OpenGL c2w with columns (right, up, -forward, eye), as built by the reference's view
    generator (interactive_segmenter.py:292-313)."""
    eye = np.asarray(eye, np.float64)
    fwd = np.asarray(target, np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, np.float64))
    right /= np.linalg.norm(right)
    true_up = np.cross(right, fwd)
    true_up /= np.linalg.norm(true_up)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, true_up, -fwd, eye
    return c2w


def projection_matrix(tanfovx: float, tanfovy: float, znear: float = 0.01, zfar: float = 100.0,
                      cx_ndc: float = 0.0, cy_ndc: float = 0.0) -> np.ndarray:
    """Perspective matrix with w_clip = z_view (z forward); cx/cy_ndc shift the principal point."""
    Pm = np.zeros((4, 4))
    Pm[0, 0] = 1.0 / tanfovx
    Pm[1, 1] = 1.0 / tanfovy
    Pm[0, 2] = cx_ndc
    Pm[1, 2] = cy_ndc
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def camera_from_c2w_opengl(c2w, fx: float, fy: float, width: int, height: int,
                           cx: float | None = None, cy: float | None = None,
                           znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """OpenGL c2w + pinhole intrinsics (pixels) -> Camera.

    Pixel convention: the reference projects to continuous image coordinates
    u = fx*x/(-z) + cx, v = H - (fy*y/(-z) + cy) (interactive_segmenter.py:1448-1456); the
    rasterizer's pixel k has its centre at coordinate k, so rasterizer_xy = (u, v) - 0.5.
    """
    c2w = np.asarray(c2w, np.float64).reshape(4, 4)
    c2w_cv = c2w.copy()
    c2w_cv[:3, 1] *= -1.0   # y up  -> y down
    c2w_cv[:3, 2] *= -1.0   # -z fwd -> +z fwd   (nerf2physic_utils.py:16)
    w2c = np.linalg.inv(c2w_cv)
    tanfovx = width / (2.0 * fx)
    tanfovy = height / (2.0 * fy)
    cx = width / 2.0 if cx is None else cx
    cy = height / 2.0 if cy is None else cy
    # principal point in NDC; image v grows downward like view-space y after the flip, but the
    # reference's cy is measured from the bottom row (v = H - ...), hence the sign on cy.
    cx_ndc = (2.0 * cx - width) / width
    cy_ndc = -(2.0 * cy - height) / height
    Pm = projection_matrix(tanfovx, tanfovy, znear, zfar, cx_ndc, cy_ndc)
    full = Pm @ w2c
    return Camera(int(height), int(width), float(tanfovx), float(tanfovy),
                  torch.tensor(w2c.T.copy(), dtype=torch.float32),
                  torch.tensor(full.T.copy(), dtype=torch.float32),
                  torch.tensor(c2w[:3, 3].copy(), dtype=torch.float32))


def camera_look_at(eye, target, up, fov_x_deg: float, width: int, height: int, **kw) -> Camera:
    """Square-pixel pinhole camera from a horizontal field of view."""
    fx = (width / 2.0) / math.tan(math.radians(fov_x_deg) / 2.0)
    return camera_from_c2w_opengl(look_at_c2w_opengl(eye, target, up), fx, fx, width, height, **kw)


def reference_six_views(center, size: float, resolution: int = 800, fov_deg: float = 50.0):
    """The six views of the reference segmenter (interactive_segmenter.py:262-313): two oblique
    (top/bottom) and four axis-aligned, at distance 2*size, fov 50 deg, square images."""
    center = np.asarray(center, np.float64)
    dist = size * 2.0
    h, d = dist, dist * 0.7
    offsets = {"top": (-d, h, 0), "bottom": (d, -h, 0), "front": (0, 0, dist),
               "back": (0, 0, -dist), "left": (-dist, 0, 0), "right": (dist, 0, 0)}
    ups = {"top": (0, 0, -1), "bottom": (0, 0, 1)}
    f = (resolution / 2.0) / math.tan(math.radians(fov_deg) / 2.0)
    cams = {}
    for name, off in offsets.items():
        c2w = look_at_c2w_opengl(center + np.asarray(off), center, ups.get(name, (0, 1, 0)))
        cams[name] = camera_from_c2w_opengl(c2w, f, f, resolution, resolution)
    return cams


def parse_transforms_json(path: str):
    """Nerfstudio transforms.json -> list of (c2w[4,4], fx, fy, cx, cy, w, h)
    (format as read by nerf2physic_utils.py:26-52; per-frame intrinsics override globals)."""
    with open(path, "r") as f:
        tr = json.load(f)
    out = []
    for fr in tr["frames"]:
        g = lambda k: fr.get(k, tr.get(k))
        out.append((np.asarray(fr["transform_matrix"], np.float64), float(g("fl_x")), float(g("fl_y")),
                    float(g("cx")), float(g("cy")), int(g("w")), int(g("h"))))
    return out


def apply_dataparser_transform(c2w, transform, scale: float) -> np.ndarray:
    """Nerfstudio dataparser transform (3x4 ``transform`` then ``scale`` on the translation), the
    pair read by nerf2physic_utils.py:55-61."""
    c2w = np.asarray(c2w, np.float64).reshape(4, 4)
    T = np.eye(4)
    T[:3, :4] = np.asarray(transform, np.float64).reshape(3, 4)
    out = T @ c2w
    out[:3, 3] *= scale
    return out


def cameras_from_nerfstudio(transforms_path: str, dataparser_path: str | None = None):
    frames = parse_transforms_json(transforms_path)
    tf, sc = None, 1.0
    if dataparser_path is not None:
        with open(dataparser_path, "r") as f:
            dp = json.load(f)
        tf, sc = dp["transform"], float(dp["scale"])
    cams = []
    for c2w, fx, fy, cx, cy, w, h in frames:
        if tf is not None:
            c2w = apply_dataparser_transform(c2w, tf, sc)
        # nerfstudio's cy is measured from the top row; convert to the bottom-origin cy used above
        cams.append(camera_from_c2w_opengl(c2w, fx, fy, w, h, cx=cx, cy=h - cy))
    return cams


def orbit_cameras(n: int, center, radius: float, fov_x_deg: float, width: int, height: int,
                  seed: int = 44):
    """Seeded inward-looking orbit/hemisphere."""
    g = np.random.default_rng(seed)
    cams = []
    for i in range(n):
        az = 2 * math.pi * (i + g.uniform(-0.25, 0.25)) / n
        el = g.uniform(-0.35, 0.35)
        eye = np.asarray(center, np.float64) + radius * np.array(
            [math.cos(az) * math.cos(el), math.sin(az) * math.cos(el), math.sin(el)])
        cams.append(camera_look_at(eye, center, (0, 0, 1), fov_x_deg, width, height))
    return cams
