"""Camera-convention known-answer tests against vectors produced by the Python reference
(tests/golden/make_camera_golden.py)."""
import json
import os

import numpy as np

from robosimgs_b200.cameras import (camera_from_c2w_opengl, look_at_c2w_opengl, reference_six_views)

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "camera_golden.json")))


def _pixel_centres(cam, pts):
    hom = np.concatenate([pts, np.ones((len(pts), 1))], 1)
    ph = hom @ cam.projmatrix.double().numpy()      # row-vector convention, transposed matrix
    ndc = ph[:, :2] / (ph[:, 3:4] + 1e-7)
    W, H = cam.image_width, cam.image_height
    return ((ndc[:, 0] + 1) * W - 1) / 2, ((ndc[:, 1] + 1) * H - 1) / 2


def test_projection_matches_reference_segmenter_and_nerfstudio_helper():
    for name, v in GOLD["views"].items():
        K = np.array(v["intrinsics"])
        W, H = v["resolution"]
        cam = camera_from_c2w_opengl(v["c2w"], K[0, 0], K[1, 1], W, H, cx=K[0, 2], cy=K[1, 2])
        px, py = _pixel_centres(cam, np.array(v["points"]))
        uv = np.array(v["uv_segmenter"])
        # rasterizer pixel k has its centre at coordinate k; the reference's continuous image
        # coordinate of that centre is k + 0.5
        assert np.abs(px + 0.5 - uv[:, 0]).max() < 1e-4, name
        assert np.abs(py + 0.5 - uv[:, 1]).max() < 1e-4, name
        uvn = np.array(v["uv_nerfstudio"])          # top-origin v; equal when cy = H/2
        assert np.abs(px + 0.5 - uvn[:, 0]).max() < 1e-4 and np.abs(py + 0.5 - uvn[:, 1]).max() < 1e-4


def test_view_space_is_opencv_and_campos_is_eye():
    v = GOLD["views"]["front"]
    K = np.array(v["intrinsics"])
    cam = camera_from_c2w_opengl(v["c2w"], K[0, 0], K[1, 1], 800, 800)
    c2w = np.array(v["c2w"])
    assert np.allclose(cam.campos.numpy(), c2w[:3, 3], atol=1e-5)
    V = cam.viewmatrix.double().numpy().T
    target = c2w[:3, 3] - c2w[:3, 2] * 3.0          # 3 units along the viewing direction (-Z of c2w)
    z = V[2, :3] @ target + V[2, 3]
    assert abs(z - 3.0) < 1e-5                       # +Z forward in view space
    assert abs(cam.tanfovx - np.tan(np.radians(25.0))) < 1e-9


def test_six_views_reproduce_reference_camera_params():
    # the committed sample run used centre/size of the openbox mesh; recover them from the fixture
    views = GOLD["views"]
    eye_f, eye_b = np.array(views["front"]["c2w"])[:3, 3], np.array(views["back"]["c2w"])[:3, 3]
    centre = (eye_f + eye_b) / 2
    size = np.linalg.norm(eye_f - eye_b) / 4.0
    cams = reference_six_views(centre, size, 800)
    for name, v in views.items():
        K = np.array(v["intrinsics"])
        ref = camera_from_c2w_opengl(v["c2w"], K[0, 0], K[1, 1], 800, 800)
        assert np.allclose(cams[name].viewmatrix.numpy(), ref.viewmatrix.numpy(), atol=2e-5), name
        assert np.allclose(cams[name].projmatrix.numpy(), ref.projmatrix.numpy(), atol=2e-4), name


def test_look_at_axes():
    c2w = look_at_c2w_opengl((0, 0, 4), (0, 0, 0), (0, 1, 0))
    assert np.allclose(c2w[:3, :3], np.eye(3)) and np.allclose(c2w[:3, 3], [0, 0, 4])


def test_hinge_fixture_is_unit_axis():
    ax = np.array(GOLD["hinge"]["axis"])
    assert abs(np.linalg.norm(ax) - 1.0) < 1e-6
    assert GOLD["joint_limits"]["upper"] == 1.57
