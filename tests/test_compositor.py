"""CPU tests of the articulated-compositor host math (hinge pose, quaternion algebra)."""
import json
import math
import os

import numpy as np

from robosimgs_b200 import compositor as cp

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "camera_golden.json")))


def test_hinge_constants_match_reference_fixture():
    assert np.allclose(cp.OPENBOX_HINGE_AXIS, GOLD["hinge"]["axis"], atol=1e-12)
    assert cp.OPENBOX_JOINT_LIMITS == (GOLD["joint_limits"]["lower"], GOLD["joint_limits"]["upper"])


def test_revolute_pose_keeps_points_on_the_axis_fixed():
    axis, origin = cp.OPENBOX_HINGE_AXIS, (0.1, -0.2, 0.3)
    T, q = cp.revolute_link_pose(axis, origin, 0.9)
    a = np.asarray(axis) / np.linalg.norm(axis)
    for s in (-1.0, 0.0, 2.5):
        p = np.asarray(origin) + s * a
        assert np.allclose(T[:, :3] @ p + T[:, 3], p, atol=1e-12)
    assert np.allclose(cp.quat_to_matrix(q), T[:, :3], atol=1e-12)
    # a point off the axis moves by the chord of the rotation
    p = np.asarray(origin) + np.array([0.5, 0, 0])
    r = np.linalg.norm(np.cross(p - origin, a))
    assert abs(np.linalg.norm(T[:, :3] @ p + T[:, 3] - p) - 2 * r * math.sin(0.45)) < 1e-9


def test_base_placement_and_scale_compose():
    bq = cp.axis_angle_quat((0, 0, 1), 0.3)
    T, q = cp.revolute_link_pose((0, 1, 0), (0, 0, 0), 0.5, base_q=bq, base_t=(1, 2, 3), scale=0.1)
    R = cp.quat_to_matrix(bq) @ cp.quat_to_matrix(cp.axis_angle_quat((0, 1, 0), 0.5))
    assert np.allclose(T[:, :3], 0.1 * R) and np.allclose(T[:, 3], (1, 2, 3))
    assert np.allclose(cp.quat_to_matrix(q), R, atol=1e-12)


def test_lid_angle_schedule_and_object_sampler():
    assert cp.lid_angle(0) == 0.0 and abs(cp.lid_angle(60) - 1.57) < 1e-12 and abs(cp.lid_angle(120)) < 1e-12
    sc, link_ids, hinge = cp.box_with_lid_gaussians(330, 170)
    assert sc.P == 500 and sc.sh_degree == 0 and (link_ids[:330] == 0).all() and (link_ids[330:] == 1).all()
    assert np.allclose(sc.rotations.norm(dim=1).numpy(), 1.0, atol=1e-5)
    assert abs(hinge[2] - 0.3) < 1e-9


# ---- the reference's own object (openbox_output/urdf): loaders vs the committed fixture ---------------------------
OPENBOX = np.load(os.path.join(os.path.dirname(__file__), "golden", "openbox_surface_samples.npz"))
REF_URDF = "/root/reference/Articulation/openbox_output/urdf"


def test_openbox_fixture_is_the_reference_object():
    """Mesh sizes as surveyed from the reference's GLBs (8416 v / 16607 f and 4410 v / 8393 f), hinge axis
    and limits of metadata.json:14-18,26-30, samples inside the mesh bounds."""
    assert OPENBOX["mesh_counts"].tolist() == [8416, 16607, 4410, 8393]
    assert np.allclose(OPENBOX["axis"], cp.OPENBOX_HINGE_AXIS, atol=1e-12)
    assert tuple(OPENBOX["limits"]) == cp.OPENBOX_JOINT_LIMITS
    assert OPENBOX["body_pts"].shape == (33_000, 3) and OPENBOX["lid_pts"].shape == (17_000, 3)
    for part in ("body", "lid"):
        pts, (lo, hi) = OPENBOX[part + "_pts"], OPENBOX[part + "_bounds"]
        assert (pts >= lo - 1e-5).all() and (pts <= hi + 1e-5).all()
        assert np.allclose(np.linalg.norm(OPENBOX[part + "_nrm"].astype(np.float64), axis=1), 1.0, atol=2e-3)
    # the hinge was moved to the origin (pipeline.py:302-307): lid and body meet next to it
    assert np.linalg.norm(OPENBOX["lid_pts"], axis=1).min() < 0.12 and np.linalg.norm(OPENBOX["body_pts"], axis=1).min() < 0.12


def test_object_from_fixture_samples():
    sc, link_ids = cp.object_from_surface_samples(OPENBOX["body_pts"], OPENBOX["body_nrm"].astype(np.float32),
                                                  OPENBOX["lid_pts"], OPENBOX["lid_nrm"].astype(np.float32))
    assert sc.P == 50_000 and int(link_ids.sum()) == 17_000 and sc.sh_degree == 0
    assert np.allclose(sc.rotations.norm(dim=1).numpy(), 1.0, atol=1e-5)
    # the splat's short axis (local +z) is the face normal
    q = sc.rotations[:200].double().numpy()
    z = np.stack([cp.quat_to_matrix(qi)[:, 2] for qi in q])
    n = OPENBOX["body_nrm"][:200].astype(np.float64)
    assert np.abs((z * n).sum(1) - 1.0).max() < 5e-3


def test_loaders_reproduce_the_fixture_from_the_reference_files():
    """Only where the reference is present (this container): GLB parser + area sampler + metadata loader give the
    committed fixture bit for bit; sampled points lie on the mesh surface."""
    import pytest
    if not os.path.isdir(REF_URDF):
        pytest.skip("reference not available on this box")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_openbox", os.path.join(os.path.dirname(__file__), "golden",
                                                                               "make_openbox_gaussians.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    d = mod.generate()
    for k in ("body_pts", "body_nrm", "lid_pts", "lid_nrm", "mesh_counts"):
        assert np.array_equal(d[k], OPENBOX[k]), k
    meta = cp.load_urdf_metadata(os.path.join(REF_URDF, "metadata.json"))
    assert np.allclose(meta["axis"], cp.OPENBOX_HINGE_AXIS) and meta["limits"] == cp.OPENBOX_JOINT_LIMITS
    v, f = cp.load_glb_mesh(os.path.join(REF_URDF, "lid_centered.glb"))
    pts, nrm = cp.gaussians_on_mesh(v, f, 500, 9)
    # every sample lies in the plane of (at least) one face with that normal: distance to the nearest face plane ~ 0
    tri = v[f]
    fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    fn /= np.maximum(np.linalg.norm(fn, axis=1, keepdims=True), 1e-20)
    d_plane = np.abs(((pts[:, None, :] - tri[None, :, 0, :]) * fn[None]).sum(-1))
    assert d_plane.min(axis=1).max() < 1e-9
    # area-uniform: the share of samples per face tracks the face areas (chi-square style bound on big faces)
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    pts2, _ = cp.gaussians_on_mesh(v, f, 200_000, 10)
    half = pts2[:, 2] > np.median(v[:, 2])
    area_half = area[(tri[:, :, 2].mean(1) > np.median(v[:, 2]))].sum() / area.sum()
    assert abs(half.mean() - area_half) < 0.02
