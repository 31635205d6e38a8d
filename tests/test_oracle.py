"""CPU tests of the oracle itself: the hand-derived analytic adjoint in
oracle/gs_oracle.c against fp64 autograd of the independent torch restatement, f32-vs-f64
consistency, and size-independent properties of the algorithm."""
import numpy as np
import pytest
import torch

from oracle import gs_oracle, gs_oracle_torch
from helpers import max_rel_err, psnr, small_scene


def _torch_grads(rs, sc, w, colors_precomp=None, cov3D=None):
    d = lambda t: None if t is None else t.double().clone().requires_grad_(True)
    m, o = d(sc.means3D), d(sc.opacities)
    s = None if colors_precomp is not None else d(sc.shs)
    cp = d(colors_precomp)
    scl, r = (None, None) if cov3D is not None else (d(sc.scales), d(sc.rotations))
    cv = d(cov3D)
    col, radii, m2d = gs_oracle_torch.render(rs, m, o, shs=s, colors_precomp=cp, scales=scl, rotations=r,
                                             cov3D_precomp=cv)
    (col * w).sum().backward()
    g = {"means3D": m.grad, "opacities": o.grad, "means2D": m2d.grad}
    for k, v in (("shs", s), ("colors_precomp", cp), ("scales", scl), ("rotations", r), ("cov3D", cv)):
        if v is not None:
            g[k] = v.grad
    return col.detach().numpy(), radii.numpy(), {k: v.numpy() for k, v in g.items()}


@pytest.mark.parametrize("degree,boost", [(3, 0.0), (1, 4.0), (0, 0.0)])
def test_analytic_backward_matches_autograd_fp64(degree, boost):
    sc, cam, rs = small_scene(P=300, degree=degree, scale_modifier=1.3, opacity_boost=boost)
    w = torch.rand(3, 56, 72, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    col, radii, tg = _torch_grads(rs, sc, w)
    st = gs_oracle.forward(rs, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations,
                           dtype=np.float64, denom_eps=0.0)
    g = gs_oracle.backward(st, w.numpy())
    assert np.abs(st.color - col).max() < 1e-12
    assert (st.radii == radii).all()
    if boost:
        assert (st.conic_opacity[:, 3] > 0.99).any()   # the 0.99 clamp is exercised
    for name, got in (("means3D", g.means3D), ("shs", g.shs), ("opacities", g.opacities), ("scales", g.scales),
                      ("rotations", g.rotations), ("means2D", g.mean2D)):
        assert max_rel_err(got, tg[name].reshape(got.shape)) < 1e-10, name


def test_analytic_backward_precomputed_inputs_fp64():
    sc, cam, rs = small_scene(P=200, degree=0)
    cols = torch.rand(200, 3, generator=torch.Generator().manual_seed(3))
    st0 = gs_oracle.forward(rs, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations,
                            dtype=np.float64)
    cov = torch.from_numpy(st0.cov3d.copy())
    cov[st0.radii <= 0] = torch.tensor([1e-3, 0, 0, 1e-3, 0, 1e-3], dtype=torch.float64)
    w = torch.rand(3, 56, 72, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    col, radii, tg = _torch_grads(rs, sc, w, colors_precomp=cols, cov3D=cov)
    st = gs_oracle.forward(rs, sc.means3D, sc.opacities, colors_precomp=cols, cov3D_precomp=cov,
                           dtype=np.float64, denom_eps=0.0)
    g = gs_oracle.backward(st, w.numpy())
    assert np.abs(st.color - col).max() < 1e-12
    assert max_rel_err(g.colors_precomp, tg["colors_precomp"]) < 1e-10
    assert max_rel_err(g.cov3D, tg["cov3D"]) < 1e-10
    assert max_rel_err(g.means3D, tg["means3D"]) < 1e-10


def test_f32_matches_f64():
    sc, cam, rs = small_scene(P=2000, degree=3, W=128, H=96)
    a = gs_oracle.forward(rs, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations,
                          dtype=np.float32)
    b = gs_oracle.forward(rs, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations,
                          dtype=np.float64)
    assert psnr(a.color, b.color) > 80.0
    assert (a.radii != b.radii).mean() < 1e-3
    w = np.random.default_rng(0).random((3, 96, 128))
    ga, gb = gs_oracle.backward(a, w), gs_oracle.backward(b, w)
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        assert max_rel_err(getattr(ga, k), getattr(gb, k)) < 1e-3, k


def test_regulariser_effect_is_small():
    """denom_eps = 1e-7 (public implementation) vs 0 (exact derivative): well inside 1e-3."""
    sc, cam, rs = small_scene(P=300, degree=1)
    w = np.random.default_rng(0).random((3, 56, 72))
    kw = dict(shs=sc.shs, scales=sc.scales, rotations=sc.rotations, dtype=np.float64)
    g0 = gs_oracle.backward(gs_oracle.forward(rs, sc.means3D, sc.opacities, denom_eps=0.0, **kw), w)
    g1 = gs_oracle.backward(gs_oracle.forward(rs, sc.means3D, sc.opacities, denom_eps=1e-7, **kw), w)
    assert max_rel_err(g1.scales, g0.scales) < 1e-4
    assert max_rel_err(g1.means3D, g0.means3D) < 1e-4


def test_binning_is_sorted_and_complete():
    sc, cam, rs = small_scene(P=1500, degree=0, W=96, H=80)
    st = gs_oracle.forward(rs, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations)
    assert st.num_rendered == int(st.tiles_touched.sum()) > 0
    covered = 0
    for t, (a, b) in enumerate(st.ranges):
        ids = st.point_list[a:b]
        d = st.depths[ids]
        assert (np.diff(d) >= 0).all()
        # ties keep index order (stable sort)
        tie = np.diff(d) == 0
        assert (np.diff(ids)[tie] > 0).all()
        covered += b - a
    assert covered == st.num_rendered


def test_empty_and_culled_scenes_render_background():
    sc, cam, rs = small_scene(P=50, degree=0, bg=(0.3, 0.6, 0.9))
    behind = sc.means3D.clone()
    behind[:, 2] += 100.0   # camera at z=2.5 looking toward -z: everything behind it
    st = gs_oracle.forward(rs, behind, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations)
    assert st.num_rendered == 0 and (st.radii == 0).all()
    assert np.allclose(st.color, np.array([0.3, 0.6, 0.9], np.float32)[:, None, None])
    st = gs_oracle.forward(rs, sc.means3D[:0], sc.opacities[:0], shs=sc.shs[:0], scales=sc.scales[:0],
                           rotations=sc.rotations[:0])
    assert np.allclose(st.color, np.array([0.3, 0.6, 0.9], np.float32)[:, None, None])


def test_colour_linearity_property():
    """With precomputed colours the image is linear in them: R(a*c1 + b*c2) = a R(c1) + b R(c2)
    (background 0).  Size-independent property used again at full size on the GPU."""
    sc, cam, rs = small_scene(P=500, degree=0, bg=(0, 0, 0))
    g = torch.Generator().manual_seed(9)
    c1, c2 = torch.rand(500, 3, generator=g, dtype=torch.float64), torch.rand(500, 3, generator=g, dtype=torch.float64)
    f = lambda c: gs_oracle.forward(rs, sc.means3D, sc.opacities, colors_precomp=c, scales=sc.scales,
                                    rotations=sc.rotations, dtype=np.float64).color
    assert np.abs(f(0.3 * c1 + 0.7 * c2) - (0.3 * f(c1) + 0.7 * f(c2))).max() < 1e-12


def test_argument_errors_match_public_interface():
    sc, cam, rs = small_scene(P=10, degree=0)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        gs_oracle.forward(rs, sc.means3D, sc.opacities, scales=sc.scales, rotations=sc.rotations)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        gs_oracle.forward(rs, sc.means3D, sc.opacities, shs=sc.shs)


def test_mark_visible():
    sc, cam, rs = small_scene(P=300, degree=0, eye=(0, 0, 0.5))
    vis = gs_oracle.mark_visible(rs, sc.means3D)
    V = rs.viewmatrix.numpy().T
    z = sc.means3D.numpy() @ V[2, :3] + V[2, 3]
    assert (vis == (z > 0.2)).all() and vis.any() and (~vis).any()
