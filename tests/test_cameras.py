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


# ---- Nerfstudio wire format (transforms.json / dataparser_transforms.json) vs the reference's readers ----------
HERE = os.path.join(os.path.dirname(__file__), "golden")
NS = json.load(open(os.path.join(HERE, "nerfstudio_golden.json")))


def _uv_top_origin(cam, pts):
    """Continuous image coordinates (u right, v DOWN from the top row, pixel k centred at k + 0.5) of world points
    through the rasterizer's own projection matrix -- the convention of nerf2physic_utils.project_3d_to_2d."""
    px, py = _pixel_centres(cam, np.asarray(pts))
    return np.stack([px + 0.5, py + 0.5], 1)


def test_parse_transforms_json_matches_reference_reader():
    from robosimgs_b200.cameras import parse_transforms_json
    fr = parse_transforms_json(os.path.join(HERE, "ns_transforms_global.json"))
    K = np.array(NS["global"]["K"])
    assert len(fr) == len(NS["global"]["c2w"]) == 5
    for (c2w, fx, fy, cx, cy, w, h), ref_c2w, ref_w2c in zip(fr, NS["global"]["c2w"], NS["global"]["w2c"]):
        assert np.array_equal(c2w, np.array(ref_c2w))
        assert np.allclose(np.linalg.inv(c2w), np.array(ref_w2c), atol=1e-12)
        assert (fx, fy, cx, cy) == (K[0, 0], K[1, 1], K[0, 2], K[1, 2]) and (w, h) == tuple(NS["image"])
    frp = parse_transforms_json(os.path.join(HERE, "ns_transforms_perframe.json"))
    for (c2w, fx, fy, cx, cy, w, h), Kp in zip(frp, NS["perframe"]["K"]):      # reference: different_Ks=True
        Kp = np.array(Kp)
        assert (fx, fy, cx, cy) == (Kp[0, 0], Kp[1, 1], Kp[0, 2], Kp[1, 2])


def test_cameras_from_nerfstudio_project_like_the_reference():
    """Cameras built from transforms.json project the reference's points where nerf2physic_utils.project_3d_to_2d puts
    them (off-centre principal point, top-origin cy, global and per-frame intrinsics)."""
    from robosimgs_b200.cameras import cameras_from_nerfstudio
    pts = np.array(NS["points_orig"])
    for key, fname in (("global", "ns_transforms_global.json"), ("perframe", "ns_transforms_perframe.json")):
        cams = cameras_from_nerfstudio(os.path.join(HERE, fname))
        for cam, uv in zip(cams, NS[key]["uv"]):
            assert (cam.image_width, cam.image_height) == tuple(NS["image"])
            assert np.abs(_uv_top_origin(cam, pts) - np.array(uv)).max() < 1e-3, key


def test_dataparser_transform_matches_the_reference_convention():
    """dataparser_transforms.json: the reference maps a Nerfstudio-space point back with inv([transform; 0 0 0 1/scale])
    (load_ns_point_cloud); cameras moved INTO Nerfstudio space by cameras_from_nerfstudio(..., dataparser) must see the
    Nerfstudio-space points exactly where the original cameras see the mapped-back points."""
    from robosimgs_b200.cameras import apply_dataparser_transform, cameras_from_nerfstudio
    dp = json.load(open(os.path.join(HERE, "ns_dataparser_transforms.json")))
    assert np.array_equal(np.array(dp["transform"]), np.array(NS["dataparser"]["transform"])) and dp["scale"] == NS["dataparser"]["scale"]
    cams = cameras_from_nerfstudio(os.path.join(HERE, "ns_transforms_global.json"), os.path.join(HERE, "ns_dataparser_transforms.json"))
    pts_ns = np.array(NS["points_ns"])
    for cam, uv in zip(cams, NS["global"]["uv"]):
        assert np.abs(_uv_top_origin(cam, pts_ns) - np.array(uv)).max() < 1e-3
    # the camera centre itself follows the same map: p_ns = scale * (R p + t)
    c2w = np.array(NS["global"]["c2w"][0])
    moved = apply_dataparser_transform(c2w, dp["transform"], dp["scale"])
    T = np.array(dp["transform"])
    assert np.allclose(moved[:3, 3], dp["scale"] * (T[:, :3] @ c2w[:3, 3] + T[:, 3]), atol=1e-12)
    assert np.allclose(moved[:3, :3], T[:, :3] @ c2w[:3, :3], atol=1e-12)
