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
