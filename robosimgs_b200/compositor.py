"""Articulated-object compositor (BASELINE config 5): background Gaussians +
object Gaussians attached to URDF links; per frame the links' rigid transforms are applied to the
object Gaussians' means and quaternions (libb200gs: b200gs_transform_gaussians) into the tail of one
persistent, concatenated parameter set, which is then rendered by the ordinary rasterizer -- the
background is never copied again.

The articulation comes from the reference's URDF pipeline: a revolute hinge with axis, origin and
limits (/root/reference/Articulation/urdf_generation/pipeline.py:290-357,
hinge_detector.py:27-110; sample values openbox_output/urdf/metadata.json:8-30, pinned in
tests/golden/camera_golden.json["hinge"]).  Object SH is restricted to degree 0 (view-independent
colour), so no SH rotation is needed.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Sequence

import numpy as np
import torch

from . import _cabi
from .scenes import SH_C0, Scene

# openbox_output/urdf/metadata.json:14-18, :26-30 (hinge axis after centring; joint limits)
OPENBOX_HINGE_AXIS = (-0.018310275957379343, -0.010006996124500128, -0.9997822732089867)
OPENBOX_JOINT_LIMITS = (0.0, 1.57)


def axis_angle_quat(axis: Sequence[float], theta: float) -> np.ndarray:
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    return np.concatenate([[math.cos(theta / 2)], math.sin(theta / 2) * a])


def quat_to_matrix(q: np.ndarray) -> np.ndarray:
    r, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                     [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                     [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def revolute_link_pose(axis, origin, theta: float, base_q=None, base_R=None, base_t=None, scale: float = 1.0):
    """World pose (3x4, quat) of a link that rotates by theta about `axis` through `origin` in the
    object frame, the object frame itself being placed by (base_R | base_q, base_t) and `scale`."""
    q = axis_angle_quat(axis, theta)
    R = quat_to_matrix(q)
    o = np.asarray(origin, np.float64)
    t = o - R @ o
    bq = np.array([1.0, 0, 0, 0]) if base_q is None else np.asarray(base_q, np.float64)
    bR = quat_to_matrix(bq) if base_R is None else np.asarray(base_R, np.float64)
    bt = np.zeros(3) if base_t is None else np.asarray(base_t, np.float64)
    Rw = scale * (bR @ R)
    tw = scale * (bR @ t) + bt
    return np.concatenate([Rw, tw[:, None]], axis=1), quat_mul(bq, q)


def lid_angle(frame: int, period: int = 120, upper: float = OPENBOX_JOINT_LIMITS[1]) -> float:
    """theta_t = upper * (0.5 - 0.5 cos(2 pi t / period))  (SURVEY 8(d), config C5)."""
    return upper * (0.5 - 0.5 * math.cos(2 * math.pi * frame / period))


def box_with_lid_gaussians(n_body: int = 33_000, n_lid: int = 17_000, seed: int = 5, size=(0.8, 0.56, 0.3)):
    """Procedural stand-in for the reference's sample object (a box whose lid hinges along one top
    edge): flat surface splats sampled uniformly by area, aligned to the face normals.  Returns
    (Scene with degree-0 colour, link_ids int32 [n], hinge origin in the object frame)."""
    g = torch.Generator().manual_seed(seed)
    sx, sy, sz = size

    def faces(lo, hi, which):
        lo, hi = torch.tensor(lo), torch.tensor(hi)
        out = []
        for axis in range(3):
            for side in (0, 1):
                if (axis, side) in which:
                    out.append((axis, side, lo, hi))
        return out

    def sample(face_list, n):
        areas = []
        for axis, side, lo, hi in face_list:
            e = hi - lo
            areas.append(float(e[(axis + 1) % 3] * e[(axis + 2) % 3]))
        probs = torch.tensor(areas) / sum(areas)
        pick = torch.multinomial(probs, n, replacement=True, generator=g)
        pts, nrm = torch.empty(n, 3), torch.zeros(n, 3)
        for fi, (axis, side, lo, hi) in enumerate(face_list):
            m = pick == fi
            k = int(m.sum())
            u = torch.rand(k, 3, generator=g) * (hi - lo) + lo
            u[:, axis] = hi[axis] if side else lo[axis]
            pts[m] = u
            nrm[m, axis] = 1.0 if side else -1.0
        return pts, nrm

    body_faces = faces((-sx / 2, -sy / 2, 0.0), (sx / 2, sy / 2, sz), {(0, 0), (0, 1), (1, 0), (1, 1), (2, 0)})
    lid_faces = faces((-sx / 2, -sy / 2, sz), (sx / 2, sy / 2, sz + 0.02), {(2, 0), (2, 1), (0, 0), (0, 1), (1, 0), (1, 1)})
    bp, bn = sample(body_faces, n_body)
    lp, ln_ = sample(lid_faces, n_lid)
    pts, nrm = torch.cat([bp, lp]), torch.cat([bn, ln_])
    n = pts.shape[0]
    # quaternion rotating +z onto the face normal
    z = torch.tensor([0.0, 0.0, 1.0]).expand(n, 3)
    w = 1.0 + (z * nrm).sum(1)
    xyz = torch.cross(z, nrm, dim=1)
    flip = w < 1e-6
    q = torch.cat([w[:, None], xyz], 1)
    q[flip] = torch.tensor([0.0, 1.0, 0.0, 0.0])
    q = q / q.norm(dim=1, keepdim=True)
    scales = torch.empty(n, 3)
    scales[:, :2] = torch.rand(n, 2, generator=g) * 0.004 + 0.004
    scales[:, 2] = 0.0008
    base = torch.where(torch.arange(n)[:, None] < n_body, torch.tensor([0.75, 0.55, 0.35]), torch.tensor([0.3, 0.45, 0.8]))
    rgb = (base + (torch.rand(n, 3, generator=g) - 0.5) * 0.15).clamp(0.02, 0.98)
    shs = ((rgb - 0.5) / SH_C0)[:, None, :]
    opac = torch.full((n, 1), 0.97)
    link_ids = torch.cat([torch.zeros(n_body, dtype=torch.int32), torch.ones(n_lid, dtype=torch.int32)])
    hinge_origin = (0.0, sy / 2, sz)          # lid hinges along the +y top edge
    return Scene(pts, shs, opac, scales, q, 0), link_ids, hinge_origin


class ArticulatedScene:
    """Concatenated (background + objects) parameter set with per-frame pose updates."""

    def __init__(self, background: Scene, objects: Scene, link_ids: torch.Tensor, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _cabi.B200GSError("b200gs needs a CUDA device; there is no CPU fallback")
        self.device = dev
        self.P_bg, self.P_obj = background.P, objects.P
        M = background.shs.shape[1]
        obj_sh = torch.zeros(objects.P, M, 3)
        obj_sh[:, :objects.shs.shape[1]] = objects.shs      # degree-0 colour, higher orders zero
        cat = lambda a, b: torch.cat([a, b]).to(dev).contiguous()
        self.means3D = cat(background.means3D, objects.means3D)
        self.shs = cat(background.shs, obj_sh)
        self.opacities = cat(background.opacities, objects.opacities)
        self.scales = cat(background.scales, objects.scales)
        self.rotations = cat(background.rotations, objects.rotations)
        self.sh_degree = background.sh_degree
        # canonical (object-frame) copies of what the pose update rewrites
        self.obj_means = objects.means3D.to(dev).contiguous()
        self.obj_rots = objects.rotations.to(dev).contiguous()
        self.obj_scales = objects.scales.to(dev).contiguous()
        self.link_ids = link_ids.to(dev, torch.int32).contiguous()
        self.n_links = int(link_ids.max().item()) + 1 if link_ids.numel() else 1

    def set_link_poses(self, transforms: np.ndarray, quats: np.ndarray, scale: float = 1.0) -> None:
        """transforms [L,3,4] (may include a uniform scale), quats [L,4] unit -> rewrite the object
        tail of means3D / rotations (and scales if scale != 1)."""
        L = _cabi.lib()
        T = torch.tensor(np.asarray(transforms, np.float32).reshape(-1, 12)).pin_memory().to(self.device, non_blocking=True)
        Q = torch.tensor(np.asarray(quats, np.float32).reshape(-1, 4)).pin_memory().to(self.device, non_blocking=True)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _cabi.check(L.b200gs_transform_gaussians(
                C.c_int32(self.P_obj), p(self.obj_means), p(self.obj_rots), p(self.link_ids), p(T), p(Q),
                C.c_int32(T.shape[0]), p(self.means3D[self.P_bg:]), p(self.rotations[self.P_bg:]),
                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        if scale != 1.0:
            self.scales[self.P_bg:] = self.obj_scales * scale

    def render(self, raster_settings):
        from .rasterizer import GaussianRasterizer
        with torch.no_grad():
            return GaussianRasterizer(raster_settings)(
                self.means3D, torch.zeros_like(self.means3D), self.opacities, shs=self.shs, scales=self.scales,
                rotations=self.rotations)


# ---- the reference's object assets: URDF metadata, GLB link meshes, surface Gaussians ------------------------------
def load_urdf_metadata(path: str) -> dict:
    """metadata.json written next to the URDF by the reference's URDF pipeline
    (/root/reference/Articulation/urdf_generation/pipeline.py:359-403; sample
    openbox_output/urdf/metadata.json): hinge axis and joint limits of the revolute joint.  The link meshes were
    translated so that the hinge passes through the origin (pipeline.py:302-307), hence origin = 0.
    Limits come from here, not from the .urdf (the committed sample's .urdf was edited afterwards)."""
    import json
    with open(path, "r") as f:
        m = json.load(f)
    axis = np.asarray(m["hinge"]["axis"], np.float64)
    return {"axis": axis / np.linalg.norm(axis), "origin": np.zeros(3),
            "limits": (float(m["joint_limits"]["lower"]), float(m["joint_limits"]["upper"])),
            "files": dict(m.get("files", {})), "original_position": np.asarray(m["hinge"]["original_position"], np.float64)}


def load_glb_mesh(path: str):
    """Triangle mesh of a binary glTF (.glb) as the reference's pipeline saves its link meshes (trimesh export:
    one buffer, POSITION + indices accessors, optional node matrices).  Hand parser -- neither trimesh nor open3d is
    a dependency of this package.  Returns (vertices float64 [V,3] with node transforms applied, faces int64 [F,3])."""
    import json
    import struct
    with open(path, "rb") as f:
        blob = f.read()
    magic, version, _ = struct.unpack_from("<III", blob, 0)
    if magic != 0x46546C67 or version != 2:
        raise ValueError(f"{path}: not a glTF 2 binary")
    off, js, bin_chunk = 12, None, None
    while off < len(blob):
        clen, ctype = struct.unpack_from("<II", blob, off)
        data = blob[off + 8:off + 8 + clen]
        if ctype == 0x4E4F534A:
            js = json.loads(data)
        elif ctype == 0x004E4942:
            bin_chunk = data
        off += 8 + clen + ((4 - clen % 4) % 4)
    if js is None or bin_chunk is None:
        raise ValueError(f"{path}: missing JSON or BIN chunk")
    comp = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
    ncomp = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}

    def accessor(i):
        a = js["accessors"][i]
        bv = js["bufferViews"][a["bufferView"]]
        dt, nc = np.dtype(comp[a["componentType"]]), ncomp[a["type"]]
        start = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * nc
        if stride == dt.itemsize * nc:
            arr = np.frombuffer(bin_chunk, dt, a["count"] * nc, start).reshape(a["count"], nc)
        else:
            arr = np.stack([np.frombuffer(bin_chunk, dt, nc, start + k * stride) for k in range(a["count"])])
        return arr

    def node_matrix(n):
        if "matrix" in n:
            return np.asarray(n["matrix"], np.float64).reshape(4, 4).T        # glTF stores column-major
        M = np.eye(4)
        if "scale" in n:
            M = np.diag(list(n["scale"]) + [1.0]) @ M
        if "rotation" in n:
            x, y, z, w = n["rotation"]
            R = np.eye(4); R[:3, :3] = quat_to_matrix(np.array([w, x, y, z])); M = R @ M
        if "translation" in n:
            T = np.eye(4); T[:3, 3] = n["translation"]; M = T @ M
        return M

    verts, faces, base = [], [], 0

    def visit(idx, parent):
        nonlocal base
        n = js["nodes"][idx]
        M = parent @ node_matrix(n)
        if "mesh" in n:
            for prim in js["meshes"][n["mesh"]]["primitives"]:
                if prim.get("mode", 4) != 4:
                    continue
                v = accessor(prim["attributes"]["POSITION"]).astype(np.float64)
                v = v @ M[:3, :3].T + M[:3, 3]
                if "indices" in prim:
                    f = accessor(prim["indices"]).astype(np.int64).reshape(-1, 3)
                else:
                    f = np.arange(len(v), dtype=np.int64).reshape(-1, 3)
                verts.append(v); faces.append(f + base); base += len(v)
        for c in n.get("children", []):
            visit(c, M)

    scene = js["scenes"][js.get("scene", 0)]
    for root in scene["nodes"]:
        visit(root, np.eye(4))
    return np.concatenate(verts), np.concatenate(faces)


def gaussians_on_mesh(vertices: np.ndarray, faces: np.ndarray, n: int, seed: int):
    """n surface samples drawn uniformly BY AREA from a triangle mesh: returns
    (points [n,3], unit face normals [n,3]); scene_from_surface_samples turns them into flat Gaussians."""
    g = np.random.default_rng(seed)
    v0, v1, v2 = (vertices[faces[:, k]] for k in range(3))
    cr = np.cross(v1 - v0, v2 - v0)
    area = 0.5 * np.linalg.norm(cr, axis=1)
    pick = g.choice(len(faces), size=n, p=area / area.sum())
    u, v = g.random(n), g.random(n)
    flip = u + v > 1.0
    u[flip], v[flip] = 1.0 - u[flip], 1.0 - v[flip]
    pts = v0[pick] + u[:, None] * (v1[pick] - v0[pick]) + v[:, None] * (v2[pick] - v0[pick])
    nrm = cr[pick] / np.maximum(np.linalg.norm(cr[pick], axis=1, keepdims=True), 1e-20)
    return pts, nrm


def scene_from_surface_samples(pts: np.ndarray, nrm: np.ndarray, seed: int, rgb=(0.7, 0.6, 0.4),
                               in_plane=(0.004, 0.008), thickness: float = 0.0008, opacity: float = 0.97) -> Scene:
    """Surface samples (+ unit normals) -> degree-0 Scene of flat splats: quaternion rotating +z onto the normal,
    in-plane scales U[in_plane], `thickness` along the normal, base colour +- 7.5 % seeded jitter."""
    g = torch.Generator().manual_seed(seed)
    n = len(pts)
    nrm_t = torch.tensor(np.asarray(nrm), dtype=torch.float32)
    z = torch.tensor([0.0, 0.0, 1.0]).expand(n, 3)
    w = 1.0 + (z * nrm_t).sum(1)
    xyz = torch.cross(z, nrm_t, dim=1)
    q = torch.cat([w[:, None], xyz], 1)
    q[w < 1e-6] = torch.tensor([0.0, 1.0, 0.0, 0.0])
    q = q / q.norm(dim=1, keepdim=True)
    scales = torch.empty(n, 3)
    scales[:, :2] = torch.rand(n, 2, generator=g) * (in_plane[1] - in_plane[0]) + in_plane[0]
    scales[:, 2] = thickness
    col = (torch.tensor(rgb, dtype=torch.float32) + (torch.rand(n, 3, generator=g) - 0.5) * 0.15).clamp(0.02, 0.98)
    return Scene(torch.tensor(np.asarray(pts), dtype=torch.float32), ((col - 0.5) / SH_C0)[:, None, :],
                 torch.full((n, 1), float(opacity)), scales, q, 0)


def articulated_object_from_urdf_dir(urdf_dir: str, n_body: int = 33_000, n_lid: int = 17_000, seed: int = 5):
    """The reference's sample object as Gaussians: body_centered.glb / lid_centered.glb named by metadata.json
    (openbox_output/urdf/), 33k + 17k surface splats.  Returns (Scene, link_ids, metadata)."""
    import os
    meta = load_urdf_metadata(os.path.join(urdf_dir, "metadata.json"))
    bv, bf = load_glb_mesh(os.path.join(urdf_dir, meta["files"].get("body_mesh", "body_centered.glb")))
    lv, lf = load_glb_mesh(os.path.join(urdf_dir, meta["files"].get("lid_mesh", "lid_centered.glb")))
    bp, bn = gaussians_on_mesh(bv, bf, n_body, seed)
    lp, ln_ = gaussians_on_mesh(lv, lf, n_lid, seed + 1)
    return object_from_surface_samples(bp, bn, lp, ln_, seed), meta


def object_from_surface_samples(body_pts, body_nrm, lid_pts, lid_nrm, seed: int = 5):
    """(Scene, link_ids) of a two-link object from surface samples of its body (link 0) and lid (link 1)."""
    body = scene_from_surface_samples(body_pts, body_nrm, seed, rgb=(0.8, 0.6, 0.4))      # URDF link colours,
    lid = scene_from_surface_samples(lid_pts, lid_nrm, seed + 1, rgb=(0.6, 0.8, 0.4))      # pipeline.py:327,335
    sc = Scene(*(torch.cat([a, b]) for a, b in zip(body[:5], lid[:5])), 0)
    link_ids = torch.cat([torch.zeros(len(body_pts), dtype=torch.int32), torch.ones(len(lid_pts), dtype=torch.int32)])
    return sc, link_ids
