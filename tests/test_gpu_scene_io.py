"""GPU tests of the data-format kernels either side of the rasterizer: .ply
activation/repack and the articulated pose update, each against a numpy statement of the same
arithmetic, then end to end through the rasterizer against the oracle."""
import math

import numpy as np
import pytest
import torch

from helpers import psnr, small_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(built):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_ply_activation_kernel_matches_numpy(tmp_path):
    from robosimgs_b200 import ply
    from robosimgs_b200.scenes import cube_scene
    for degree, P in ((3, 1000), (0, 333), (1, 129)):
        sc, _ = cube_scene(P=P, seed=11, degree=degree)
        sc.rotations.mul_(torch.rand(P, 1, generator=torch.Generator().manual_seed(1)) + 0.5)   # un-normalised
        path = str(tmp_path / f"s{degree}.ply")
        ply.write_gaussian_ply(path, sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations)
        v, lay = ply.read_gaussian_ply(path)
        got = ply.load_gaussian_ply(path, "cuda:0")
        v64 = v.astype(np.float64)
        M = 1 + lay["n_rest"]
        assert got.sh_degree == degree and got.shs.shape == (P, M, 3)
        assert np.array_equal(got.means3D.cpu().numpy(), v[:, 0:3])
        shs = np.concatenate([v[:, lay["off_fdc"]:lay["off_fdc"] + 3][:, None, :],
                              v[:, lay["off_frest"]:lay["off_frest"] + 3 * lay["n_rest"]].reshape(P, 3, lay["n_rest"]).transpose(0, 2, 1)], 1)
        assert np.array_equal(got.shs.cpu().numpy(), shs)
        assert np.allclose(got.opacities.cpu().numpy()[:, 0], 1 / (1 + np.exp(-v64[:, lay["off_opacity"]])), rtol=2e-6)
        assert np.allclose(got.scales.cpu().numpy(), np.exp(v64[:, lay["off_scale"]:lay["off_scale"] + 3]), rtol=2e-6)
        q = v64[:, lay["off_rot"]:lay["off_rot"] + 4]
        assert np.allclose(got.rotations.cpu().numpy(), q / np.linalg.norm(q, axis=1, keepdims=True), atol=1e-6)


def test_config2_through_ply_matches_oracle(tmp_path):
    """BASELINE config 2 data path, reduced: tabletop scene -> .ply -> fused activation kernel ->
    render of a reference view, against the oracle on the original tensors."""
    from oracle import gs_oracle
    from robosimgs_b200 import GaussianRasterizer, ply
    from robosimgs_b200.scenes import settings_from_camera, tabletop_scene
    sc, cams = tabletop_scene(P=20_000, resolution=320)
    path = str(tmp_path / "tabletop.ply")
    ply.write_gaussian_ply(path, sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations)
    g = ply.load_gaussian_ply(path, "cuda:0")
    cam = cams["front"]
    rs = settings_from_camera(cam, 3, device="cuda:0")
    with torch.no_grad():
        color, _ = GaussianRasterizer(rs)(g.means3D, torch.zeros_like(g.means3D), g.opacities, shs=g.shs,
                                          scales=g.scales, rotations=g.rotations)
    st = gs_oracle.forward(settings_from_camera(cam, 3), sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales,
                           rotations=sc.rotations, dtype=np.float64)
    assert psnr(color.cpu().numpy(), st.color) >= 60.0


def test_transform_kernel_and_articulated_composite_match_oracle():
    """BASELINE config 5, reduced: background + box-with-lid object, lid posed about the reference
    hinge axis; pose kernel vs numpy, composite frame vs the oracle on numpy-posed parameters."""
    from oracle import gs_oracle
    from robosimgs_b200 import compositor as cp
    from robosimgs_b200.scenes import Scene, settings_from_camera
    bg, cam, rs = small_scene(P=1500, degree=1, W=200, H=152)
    obj, link_ids, hinge = cp.box_with_lid_gaussians(900, 500, seed=5)
    art = cp.ArticulatedScene(bg, obj, link_ids, "cuda:0")
    base_q = cp.axis_angle_quat((1, 0.3, 0), 0.7)
    scale, base_t = 1.5, (-0.2, -0.3, 0.4)
    for theta in (0.0, cp.lid_angle(37)):
        T0, q0 = cp.revolute_link_pose(cp.OPENBOX_HINGE_AXIS, hinge, 0.0, base_q=base_q, base_t=base_t, scale=scale)
        T1, q1 = cp.revolute_link_pose((1, 0, 0), hinge, theta, base_q=base_q, base_t=base_t, scale=scale)
        art.set_link_poses(np.stack([T0, T1]), np.stack([q0, q1]), scale=scale)
        # numpy statement of the same pose update
        T = np.stack([T0, T1])[link_ids.numpy()]
        Q = np.stack([q0, q1])[link_ids.numpy()]
        m = np.einsum("nij,nj->ni", T[:, :, :3], obj.means3D.numpy().astype(np.float64)) + T[:, :, 3]
        r = np.stack([cp.quat_mul(Q[i], obj.rotations.numpy()[i].astype(np.float64)) for i in range(obj.P)])
        assert np.allclose(art.means3D[art.P_bg:].cpu().numpy(), m, atol=2e-6)
        assert np.allclose(art.rotations[art.P_bg:].cpu().numpy(), r, atol=2e-6)
        assert torch.equal(art.means3D[:art.P_bg].cpu(), bg.means3D)            # background untouched
        color, radii = art.render(settings_from_camera(cam, 1, bg=(0.2, 0.1, 0.4), device="cuda:0"))
        obj_sh = np.zeros((obj.P, 4, 3), np.float32); obj_sh[:, :1] = obj.shs.numpy()
        st = gs_oracle.forward(rs, np.concatenate([bg.means3D.numpy(), m]),
                               np.concatenate([bg.opacities.numpy(), obj.opacities.numpy()]),
                               shs=np.concatenate([bg.shs.numpy(), obj_sh]),
                               scales=np.concatenate([bg.scales.numpy(), obj.scales.numpy() * scale]),
                               rotations=np.concatenate([bg.rotations.numpy(), r]), dtype=np.float64)
        assert psnr(color.cpu().numpy(), st.color) >= 60.0
        assert (radii[art.P_bg:] > 0).sum() > 100        # the object is in view


def test_fused_photometric_loss_matches_torch():
    from robosimgs_b200.losses import mse_loss, photometric_loss
    g = torch.Generator().manual_seed(8)
    a = torch.rand(3, 120, 160, generator=g).cuda().requires_grad_(True)
    b = torch.rand(3, 120, 160, generator=g).cuda()
    for fused, ref in ((lambda: mse_loss(a, b), lambda: ((a - b) ** 2).mean()),
                       (lambda: photometric_loss(a, b, 0.3, 0.8), lambda: (0.3 * (a - b) ** 2 + 0.8 * (a - b).abs()).mean())):
        a.grad = None
        l1 = fused(); (l1 * 1.7).backward(); g1 = a.grad.clone()
        a.grad = None
        l2 = ref(); (l2 * 1.7).backward(); g2 = a.grad.clone()
        assert abs(float(l1) - float(l2)) < 1e-6 * max(1.0, abs(float(l2)))
        assert torch.allclose(g1, g2, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("graphs", [True, False, "node-priority"])
def test_scene_renderer_host_frames_match_plain_path(graphs):
    """sweep.SceneRenderer (host camera in, pinned 8-bit frame out, deferred pair check, one CUDA graph
    per frame slot) returns exactly the frames of forward + export_rgb8, also across a view change that
    overflows the captured pair capacity (frame rendered again, graphs re-captured).  "node-priority": the frame
    graphs are instantiated by the library (b200gs_graph_instantiate: binning chain high, compositing low)."""
    prio = 1 if graphs == "node-priority" else 0
    graphs = bool(graphs)
    from robosimgs_b200 import GaussianRasterizer, export_rgb8
    from robosimgs_b200.cameras import camera_look_at, orbit_cameras
    from robosimgs_b200.scenes import cube_scene, settings_from_camera
    from robosimgs_b200.sweep import SceneRenderer
    dev = torch.device("cuda:0")
    sc, _ = cube_scene(P=60000, seed=9, degree=1)
    scene = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    far = orbit_cameras(9, (0, 0, 0), 9.0, 50.0, 320, 240, seed=3)            # few pairs
    near = [camera_look_at((0.2, 0.1, 2.2), (0, 0, 0), (0, 1, 0), 50.0, 320, 240)] * 3   # many more pairs
    cams = far + near + far[:4]
    from robosimgs_b200 import _cabi
    _cabi.set_option("bin_shift", 0)           # 16-px bins: the near view has several times the pairs of the far ones
    try:
        r = SceneRenderer(scene, 1, bg, 240, 320, streams=2, graphs=graphs, node_priority=prio)
        m2 = torch.zeros_like(scene["means3D"])
        got, handles = [], []
        for cam in cams:
            while len(handles) >= r.in_flight_limit():
                got.append(r.collect(handles.pop(0)).clone())
            handles.append(r.submit(cam))
        while handles:
            got.append(r.collect(handles.pop(0)).clone())
    finally:
        _cabi.set_option("bin_shift", -1)
    assert r.redone >= 1                       # the jump to the near camera overflowed the capacity
    if prio:
        assert r.priority_nodes[0] >= 1 and r.priority_nodes[1] >= 4      # compositing low, the chain high
        r.close()
    with torch.no_grad():
        for cam, frame in zip(cams, got):
            rs = settings_from_camera(cam, 1, bg=(0.1, 0.2, 0.3), device=dev)
            col, _ = GaussianRasterizer(rs)(scene["means3D"], m2, scene["opacities"], shs=scene["shs"],
                                            scales=scene["scales"], rotations=scene["rotations"])
            assert torch.equal(export_rgb8(col).cpu(), frame)


def test_config5_openbox_composite_matches_oracle():
    """BASELINE config C5 with the REFERENCE's object: surface Gaussians of openbox_output/urdf/body_centered.glb and
    lid_centered.glb (tests/golden/openbox_surface_samples.npz, generated from the reference files), scaled 0.1 into
    the scene, the lid rotated about the fixture's hinge axis (metadata.json:14-18) through the origin at the C5
    schedule's angles theta_t != 0; pose kernel vs numpy, composite frame vs the fp64 oracle on numpy-posed parameters."""
    import os
    from oracle import gs_oracle
    from robosimgs_b200 import compositor as cp
    from robosimgs_b200.scenes import settings_from_camera
    ob = np.load(os.path.join(os.path.dirname(__file__), "golden", "openbox_surface_samples.npz"))
    sel_b, sel_l = slice(0, 33_000, 4), slice(0, 17_000, 4)              # a quarter of the samples keeps the oracle quick
    obj, link_ids = cp.object_from_surface_samples(ob["body_pts"][sel_b], ob["body_nrm"][sel_b].astype(np.float32),
                                                   ob["lid_pts"][sel_l], ob["lid_nrm"][sel_l].astype(np.float32))
    bg, cam, rs = small_scene(P=2500, degree=1, W=256, H=192, eye=(0.45, 0.35, 0.55), fov=55.0)
    art = cp.ArticulatedScene(bg, obj, link_ids, "cuda:0")
    axis, scale = ob["axis"], 0.1
    base_q, base_t = cp.axis_angle_quat((1, 0, 0), -math.pi / 2), (0.05, -0.02, 0.0)     # hinge axis (~ -z) upright-ish
    lid_moved = []
    for frame in (0, 17, 60):
        theta = cp.lid_angle(frame)
        T0, q0 = cp.revolute_link_pose(axis, (0, 0, 0), 0.0, base_q=base_q, base_t=base_t, scale=scale)
        T1, q1 = cp.revolute_link_pose(axis, (0, 0, 0), theta, base_q=base_q, base_t=base_t, scale=scale)
        art.set_link_poses(np.stack([T0, T1]), np.stack([q0, q1]), scale=scale)
        T = np.stack([T0, T1])[link_ids.numpy()]
        Q = np.stack([q0, q1])[link_ids.numpy()]
        m = np.einsum("nij,nj->ni", T[:, :, :3], obj.means3D.numpy().astype(np.float64)) + T[:, :, 3]
        r = np.stack([cp.quat_mul(Q[i], obj.rotations.numpy()[i].astype(np.float64)) for i in range(obj.P)])
        assert np.allclose(art.means3D[art.P_bg:].cpu().numpy(), m, atol=2e-6)
        assert np.allclose(art.rotations[art.P_bg:].cpu().numpy(), r, atol=2e-6)
        lid_moved.append(m[link_ids.numpy() == 1].copy())
        color, radii = art.render(settings_from_camera(cam, 1, bg=(0.2, 0.1, 0.4), device="cuda:0"))
        obj_sh = np.zeros((obj.P, 4, 3), np.float32); obj_sh[:, :1] = obj.shs.numpy()
        st = gs_oracle.forward(rs, np.concatenate([bg.means3D.numpy(), m]),
                               np.concatenate([bg.opacities.numpy(), obj.opacities.numpy()]),
                               shs=np.concatenate([bg.shs.numpy(), obj_sh]),
                               scales=np.concatenate([bg.scales.numpy(), obj.scales.numpy() * scale]),
                               rotations=np.concatenate([bg.rotations.numpy(), r]), dtype=np.float64)
        assert psnr(color.cpu().numpy(), st.color) >= 60.0, frame
        assert (radii[art.P_bg:] > 0).sum() > 1000, frame            # the object is in view
    # the lid really swings: at the top of the schedule (1.57 rad) its points moved by up to ~ the lid's extent
    assert np.linalg.norm(lid_moved[2] - lid_moved[0], axis=1).max() > 0.1
    # points on the hinge axis stay put: the lid sample closest to the axis moves least
    d_axis = np.linalg.norm(np.cross(ob["lid_pts"][sel_l].astype(np.float64), axis), axis=1)
    k = int(np.argmin(d_axis))
    assert np.linalg.norm(lid_moved[2][k] - lid_moved[0][k]) < 2.0 * scale * d_axis[k] + 1e-6
