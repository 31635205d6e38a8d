"""GPU parity tests: the CUDA path (through the public operator -> ctypes -> C ABI of
libb200gs.so) against the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): image PSNR >= 60 dB vs the oracle render; gradient
max-rel-err < 1e-3, defined as max|got - ref| / max|ref| per gradient tensor, ref = fp64 oracle.
"""
import numpy as np
import pytest
import torch

from helpers import gpu_render, max_rel_err, psnr, small_scene

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-3
PSNR_MIN = 60.0


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(built):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from robosimgs_b200 import _cabi
    _cabi.lib()   # fail loudly if the extension is missing


@pytest.fixture(autouse=True, params=["auto", "4px"])
def _compositing_kernels(request):
    """Every test runs twice: with the automatic kernel choice (one pixel per thread on these small images)
    and with the four-pixels-per-thread kernels forced (what 1080p frames use)."""
    from robosimgs_b200 import _cabi
    _cabi.set_option("render", -1 if request.param == "auto" else 1)
    yield
    _cabi.set_option("render", -1)


def _oracle(rs, sc, dtype=np.float64, **kw):
    from oracle import gs_oracle
    if not kw:
        kw = dict(shs=sc.shs, scales=sc.scales, rotations=sc.rotations)
    return gs_oracle.forward(rs, sc.means3D, sc.opacities, dtype=dtype, **kw)


def _check_grads(grads, ref, names):
    for n in names:
        r = getattr(ref, n)
        err = max_rel_err(grads[n].reshape(r.shape), r)
        assert err < GRAD_TOL, f"{n}: max-rel-err {err:.3e}"


def test_config1_cube_forward_matches_oracle():
    """BASELINE config 1: 10k-Gaussian cube, 256x256, SH degree 0."""
    from robosimgs_b200.scenes import cube_scene, settings_from_camera
    sc, cam = cube_scene()
    rs = settings_from_camera(cam, 0)
    color, radii, _ = gpu_render(sc, cam, 0)
    for dt in (np.float32, np.float64):
        st = _oracle(rs, sc, dt)
        assert psnr(color, st.color) >= PSNR_MIN
        assert (radii != st.radii).mean() <= 1e-3          # ceil() flips are 1-ulp events
        assert ((radii > 0) == (st.radii > 0)).mean() >= 0.999
    assert psnr(color, st.color) > 90.0                     # in practice far above the bar


@pytest.mark.parametrize("degree,bg,mod,boost", [(3, (0.2, 0.1, 0.4), 1.0, 0.0), (2, (0, 0, 0), 1.3, 0.0),
                                                 (1, (1, 1, 1), 1.0, 4.0), (0, (0, 0, 0), 0.7, 0.0)])
def test_forward_backward_matches_fp64_oracle(degree, bg, mod, boost):
    from oracle import gs_oracle
    sc, cam, rs = small_scene(P=3000, degree=degree, W=200, H=136, bg=bg, scale_modifier=mod,
                              opacity_boost=boost)
    w = torch.rand(3, 136, 200, generator=torch.Generator().manual_seed(11))
    color, radii, grads = gpu_render(sc, cam, degree, bg=bg, scale_modifier=mod, grad_weight=w)
    st = _oracle(rs, sc)
    assert psnr(color, st.color) >= PSNR_MIN
    assert (radii != st.radii).mean() <= 1e-3
    ref = gs_oracle.backward(st, w.numpy())
    _check_grads(grads, ref, ("means3D", "shs", "opacities", "scales", "rotations", "means2D"))


def test_precomputed_colour_and_covariance_path():
    from oracle import gs_oracle
    sc, cam, rs = small_scene(P=2000, degree=0, W=160, H=120)
    cols = torch.rand(2000, 3, generator=torch.Generator().manual_seed(3))
    st0 = _oracle(rs, sc, np.float32)
    cov = torch.from_numpy(st0.cov3d.copy())
    cov[st0.radii <= 0] = torch.tensor([1e-3, 0, 0, 1e-3, 0, 1e-3])
    w = torch.rand(3, 120, 160, generator=torch.Generator().manual_seed(12))
    color, radii, grads = gpu_render(sc, cam, 0, bg=(0.2, 0.1, 0.4), grad_weight=w, colors_precomp=cols,
                                     cov3D_precomp=cov)
    st = _oracle(rs, sc, colors_precomp=cols, cov3D_precomp=cov)
    assert psnr(color, st.color) >= PSNR_MIN
    ref = gs_oracle.backward(st, w.numpy())
    ref.cov3D_precomp = ref.cov3D
    _check_grads(grads, ref, ("means3D", "colors_precomp", "opacities", "cov3D_precomp", "means2D"))


def test_tile_culling_only_drops_non_contributing_pairs():
    """The tight per-row ellipse spans must never change the image: compare against the oracle
    (full 3-sigma rect) on a scene of very anisotropic, very large and very faint splats."""
    sc, cam, rs = small_scene(P=1500, degree=0, W=256, H=192, big=0)
    g = torch.Generator().manual_seed(4)
    sc.scales[:500, 0] *= 30.0                     # needles
    sc.scales[500:520] *= 40.0                     # screen-filling blobs
    sc.opacities[520:900] = torch.rand(380, 1, generator=g) * 0.02   # around the 1/255 threshold
    color, radii, _ = gpu_render(sc, cam, 0, bg=(0.2, 0.1, 0.4))
    st = _oracle(rs, sc)
    assert psnr(color, st.color) >= 80.0
    # fp32-vs-fp64 flips of the alpha >= 1/255 test at splat borders change a pixel by at most
    # alpha*T*c ~ 0.004; a wrongly dropped tile would show up as many larger errors
    diff = np.abs(color - st.color)
    assert diff.max() < 4.5e-3
    assert (diff > 5e-4).mean() < 1e-3


def test_num_rendered_not_larger_than_reference_rects():
    from robosimgs_b200 import GaussianRasterizer, rasterizer
    from robosimgs_b200.scenes import settings_from_camera
    sc, cam, rs = small_scene(P=3000, degree=0, W=256, H=192)
    st = _oracle(rs, sc, np.float32)
    dev = torch.device("cuda:0")
    rs_gpu = settings_from_camera(cam, 0, bg=(0.2, 0.1, 0.4), device=dev)
    m = sc.means3D.to(dev).requires_grad_(True)
    color, _ = GaussianRasterizer(rs_gpu)(m, torch.zeros_like(m), sc.opacities.to(dev), shs=sc.shs.to(dev),
                                          scales=sc.scales.to(dev), rotations=sc.rotations.to(dev))
    D = color.grad_fn.num_rendered
    assert 0 < D <= st.num_rendered


def test_edge_cases_empty_culled_and_ragged_image():
    from robosimgs_b200.scenes import Scene
    sc, cam, rs = small_scene(P=64, degree=0, W=37, H=23, bg=(0.3, 0.6, 0.9))   # not a multiple of 16
    color, radii, _ = gpu_render(sc, cam, 0, bg=(0.3, 0.6, 0.9))
    st = _oracle(rs, sc)
    assert color.shape == (3, 23, 37) and psnr(color, st.color) >= PSNR_MIN
    # everything behind the camera
    behind = Scene(sc.means3D + torch.tensor([0, 0, 100.0]), sc.shs, sc.opacities, sc.scales, sc.rotations, 0)
    w = torch.ones(3, 23, 37)
    color, radii, grads = gpu_render(behind, cam, 0, bg=(0.3, 0.6, 0.9), grad_weight=w)
    assert (radii == 0).all()
    assert np.allclose(color, np.array([0.3, 0.6, 0.9], np.float32)[:, None, None])
    assert all(np.all(g == 0) for g in grads.values())
    # zero Gaussians
    empty = Scene(sc.means3D[:0], sc.shs[:0], sc.opacities[:0], sc.scales[:0], sc.rotations[:0], 0)
    color, radii, _ = gpu_render(empty, cam, 0, bg=(0.3, 0.6, 0.9))
    assert radii.shape == (0,)
    assert np.allclose(color, np.array([0.3, 0.6, 0.9], np.float32)[:, None, None])


def test_long_tile_lists_cross_chunk_boundaries():
    """> 256 and > 512 splats in one tile exercise the 2-stage TMA ring (refill + early drain)."""
    from oracle import gs_oracle
    sc, cam, rs = small_scene(P=6000, degree=0, W=64, H=48, big=0, fov=35.0)
    sc.opacities.mul_(0.15)        # keep transmittance alive through long lists
    w = torch.rand(3, 48, 64, generator=torch.Generator().manual_seed(13))
    color, radii, grads = gpu_render(sc, cam, 0, bg=(0.2, 0.1, 0.4), grad_weight=w)
    st = _oracle(rs, sc)
    assert (st.ranges[:, 1] - st.ranges[:, 0]).max() > 600
    assert psnr(color, st.color) >= PSNR_MIN
    ref = gs_oracle.backward(st, w.numpy())
    _check_grads(grads, ref, ("means3D", "shs", "opacities", "scales", "rotations"))
    # opaque variant: pixels saturate early -> CTA-level early-out with a prefetched chunk in flight
    sc.opacities.fill_(0.95)
    color, _, _ = gpu_render(sc, cam, 0, bg=(0.2, 0.1, 0.4))
    assert psnr(color, _oracle(rs, sc).color) >= PSNR_MIN


def test_debug_mode_and_mark_visible():
    from oracle import gs_oracle
    from robosimgs_b200 import GaussianRasterizer
    from robosimgs_b200.scenes import settings_from_camera
    sc, cam, rs = small_scene(P=500, degree=1, eye=(0, 0, 0.5))
    color, _, _ = gpu_render(sc, cam, 1, bg=(0.2, 0.1, 0.4), debug=True)
    assert psnr(color, _oracle(rs, sc).color) >= PSNR_MIN
    dev = torch.device("cuda:0")
    vis = GaussianRasterizer(settings_from_camera(cam, 1, device=dev)).markVisible(sc.means3D.to(dev))
    assert vis.dtype == torch.bool
    assert (vis.cpu().numpy() == gs_oracle.mark_visible(rs, sc.means3D)).all()


def test_forward_is_deterministic_and_linear_in_colours():
    """Size-independent properties on a larger scene (50k splats, 640x480): bit-identical repeat
    renders, and linearity of the image in precomputed colours."""
    from robosimgs_b200.scenes import cube_scene
    from robosimgs_b200.cameras import camera_look_at
    sc, _ = cube_scene(P=50_000, seed=21, degree=0)
    cam = camera_look_at((0.4, 0.2, 3.0), (0, 0, 0), (0, 1, 0), 55.0, 640, 480)
    g = torch.Generator().manual_seed(9)
    c1, c2 = torch.rand(50_000, 3, generator=g), torch.rand(50_000, 3, generator=g)
    f = lambda c: gpu_render(sc, cam, 0, colors_precomp=c)[0]
    a, b = f(c1), f(c1)
    assert np.array_equal(a, b)
    mix = f(0.25 * c1 + 0.75 * c2)
    assert np.abs(mix - (0.25 * a + 0.75 * f(c2))).max() < 1e-5


_C2 = {}


def _c2_scene():
    if not _C2:
        from robosimgs_b200.scenes import tabletop_scene
        _C2["scene"], _C2["cams"] = tabletop_scene()
        _C2["oracle"] = {}
    return _C2


@pytest.mark.parametrize("view", ["top", "bottom", "front", "back", "left", "right"])
def test_config2_tabletop_view_matches_oracle(view):
    """BASELINE config 2: 200k-Gaussian tabletop, 800x800, SH degree 3, forward only -- all six reference views
    (interactive_segmenter.py:262-273).  'front'/'back' look straight down at the slab (140k splats within 1 % of
    one depth: the worst case of any depth-sliced binning), 'left'/'right' see it edge-on."""
    from robosimgs_b200.scenes import settings_from_camera
    c2 = _c2_scene()
    sc, cam = c2["scene"], c2["cams"][view]
    color, radii, _ = gpu_render(sc, cam, 3)
    if view not in c2["oracle"]:
        c2["oracle"][view] = _oracle(settings_from_camera(cam, 3), sc, np.float32)
    st = c2["oracle"][view]
    assert psnr(color, st.color) >= PSNR_MIN
    assert (radii != st.radii).mean() <= 1e-3


def test_speculative_pair_capacity_paths_are_bit_identical():
    """pair_capacity_hint: exact (sync) mode, a generous hint (padded sort) and a too-small hint
    (binning stage redone) must all give bit-identical images and the same pair count."""
    from robosimgs_b200 import GaussianRasterizer, rasterizer
    from robosimgs_b200.scenes import settings_from_camera
    sc, cam, rs = small_scene(P=4000, degree=1, W=320, H=240)
    dev = torch.device("cuda:0")
    rs_gpu = settings_from_camera(cam, 1, bg=(0.2, 0.1, 0.4), device=dev)
    args = [sc.means3D.to(dev).requires_grad_(True), torch.zeros(4000, 3, device=dev), sc.opacities.to(dev)]
    kw = dict(shs=sc.shs.to(dev), scales=sc.scales.to(dev), rotations=sc.rotations.to(dev))
    key = (dev.index, 4000, 240, 320)
    outs = []
    rasterizer.ADAPT_BIN_SIZE = False                       # keep the bin size (and so the pair count) fixed
    try:
        for hint in (0, 10_000_000, 17):
            rasterizer._PAIR_HINTS.pop(key, None)
            if hint:
                rasterizer._PAIR_HINTS[key] = hint
            color, radii = GaussianRasterizer(rs_gpu)(*args, **kw)
            outs.append((color.detach().cpu().numpy(), color.grad_fn.num_rendered))
            assert rasterizer._PAIR_HINTS[key] >= color.grad_fn.num_rendered
            w = torch.ones_like(color)
            (color * w).sum().backward()                       # backward works from every path
            assert torch.isfinite(args[0].grad).all()
    finally:
        rasterizer.ADAPT_BIN_SIZE = True
    assert outs[0][1] == outs[1][1] == outs[2][1] > 17
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][0], outs[2][0])


def test_deferred_pair_check_ticket_protocol():
    """forward_deferred never blocks on the pair count: same image as forward when the hint suffices
    (ticket ok), ticket not ok when the speculative capacity was too small (caller renders again)."""
    from robosimgs_b200 import GaussianRasterizer, _cabi, rasterizer
    from robosimgs_b200.scenes import settings_from_camera
    dev = torch.device("cuda:0")
    sc, cam, _ = small_scene(P=30000, degree=1, W=400, H=300)
    rs = settings_from_camera(cam, 1, bg=(0.2, 0.1, 0.4), device=dev)
    t = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    m2 = torch.zeros_like(t["means3D"])
    kw = dict(shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
    try:
        _cabi.set_option("bin_shift", 0)
        rasterizer.ADAPT_BIN_SIZE = False
        with torch.no_grad():
            r = GaussianRasterizer(rs)
            rasterizer._PAIR_HINTS.clear()
            c0, r0, t0 = r.forward_deferred(t["means3D"], m2, t["opacities"], **kw)     # no hint yet: exact path
            assert t0.ok()
            ref, rref = r(t["means3D"], m2, t["opacities"], **kw)
            assert torch.equal(c0, ref)
            c1, r1, t1 = r.forward_deferred(t["means3D"], m2, t["opacities"], **kw)     # hint from the last frames
            assert t1.hint > 0 and t1.ok()
            assert torch.equal(c1, ref) and torch.equal(r1, rref)
            key = t1.hint_key
            D = rasterizer._PAIR_HINTS[key]
            assert D > 40000, D                                     # above the 32768 slots of head-room
            rasterizer._PAIR_HINTS[key] = 16                        # a stale, far too small hint
            c2, _, t2 = r.forward_deferred(t["means3D"], m2, t["opacities"], **kw)
            assert not t2.ok()                                      # incomplete frame: must be rendered again
            assert rasterizer._PAIR_HINTS[key] >= D                 # ... and the hint has recovered
            c3, _, t3 = r.forward_deferred(t["means3D"], m2, t["opacities"], **kw)
            assert t3.ok() and torch.equal(c3, ref)
    finally:
        _cabi.set_option("bin_shift", -1)
        rasterizer.ADAPT_BIN_SIZE = True


@pytest.mark.parametrize("W,H", [(1280, 1024), (1282, 1022), (200, 120), (201, 119)])
def test_fused_rgb8_frame_equals_export_of_the_fp32_frame(W, H):
    """B200GS_OUT_RGB8: the compositing kernel writes the [H][W][3] 8-bit frame itself -- bit for bit what
    b200gs_export_rgb8 makes of the fp32 frame.  Both compositing kernels (four pixels per thread from 4096 tiles up,
    one pixel per thread below), with and without the 128-bit path (W % 4), synchronous (no hint yet) and deferred."""
    from robosimgs_b200 import GaussianRasterizer, export_rgb8, rasterizer
    from robosimgs_b200.scenes import settings_from_camera
    dev = torch.device("cuda:0")
    sc, cam, _ = small_scene(P=6000, degree=2, W=W, H=H)
    rs = settings_from_camera(cam, 2, bg=(0.2, 0.6, 0.4), device=dev)
    t = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    m2 = torch.zeros_like(t["means3D"])
    kw = dict(shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
    with torch.no_grad():
        r = GaussianRasterizer(rs)
        rasterizer._PAIR_HINTS.clear()
        ref, rref = r(t["means3D"], m2, t["opacities"], **kw)
        want = export_rgb8(ref)
        assert want.float().std() > 10                           # a real picture, not a constant
        rasterizer._PAIR_HINTS.clear()
        for attempt in range(2):                                 # first call: exact path; second: deferred
            out = torch.full((H, W, 3), 77, dtype=torch.uint8, device=dev)
            got, radii, ticket = r.forward_deferred(t["means3D"], m2, t["opacities"],
                                                    options=rasterizer.DeferOptions(rgb8=out), **kw)
            assert ticket.ok() and (ticket.hint > 0) == (attempt == 1)
            assert got.dtype == torch.uint8 and got.data_ptr() == out.data_ptr()
            assert torch.equal(out, want) and torch.equal(radii, rref)
        with pytest.raises(Exception, match="rgb8"):
            r.forward_deferred(t["means3D"], m2, t["opacities"],
                               options=rasterizer.DeferOptions(rgb8=torch.empty((H, W, 4), dtype=torch.uint8, device=dev)), **kw)


def test_training_step_with_deferred_pair_check_matches_oracle_and_recovers_from_overflow():
    """GaussianRasterizer.defer_pair_check (opt-in): forward returns without waiting for the pair count, backward
    launches the adjoint and only then validates.  (a) same image, gradients within the usual bound of the fp64
    oracle; (b) with a stale, far too small capacity hint backward raises PairCapacityExceeded BEFORE any gradient
    reaches a leaf, and train.backward_or_retry renders the camera again -- same gradients as the blocking path."""
    from oracle import gs_oracle
    from robosimgs_b200 import GaussianRasterizer, PairCapacityExceeded, _cabi, rasterizer
    from robosimgs_b200.scenes import settings_from_camera
    from robosimgs_b200.train import backward_or_retry
    dev = torch.device("cuda:0")
    sc, cam, rs_cpu = small_scene(P=30000, degree=1, W=400, H=300)
    rs = settings_from_camera(cam, 1, bg=(0.2, 0.1, 0.4), device=dev)
    names = ("means3D", "shs", "opacities", "scales", "rotations")
    leaves = {k: getattr(sc, k).to(dev).clone().requires_grad_(True) for k in names}
    m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
    w = torch.rand(3, 300, 400, generator=torch.Generator().manual_seed(77))
    wd = w.to(dev)
    ref = gs_oracle.backward(_oracle(rs_cpu, sc), w.numpy())
    r = GaussianRasterizer(rs)
    frames = []

    def loss_fn():
        color, _ = r(leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
                     rotations=leaves["rotations"])
        frames.append(color.detach())
        return (color * wd).sum()

    def check_grads():
        for k in names:
            got = leaves[k].grad.cpu().numpy().reshape(getattr(ref, k).shape)
            assert max_rel_err(got, getattr(ref, k)) < GRAD_TOL, k
            leaves[k].grad = None
    try:
        _cabi.set_option("bin_shift", 0)
        rasterizer.ADAPT_BIN_SIZE = False
        rasterizer._PAIR_HINTS.clear()
        loss_fn().backward()                                    # blocking path: establishes the hint
        check_grads()
        r.defer_pair_check = True
        loss_fn().backward()                                    # deferred and validated by backward
        assert r.last_ticket is not None and r.last_ticket.ok() and r.last_ticket.pairs > 40000
        assert torch.equal(frames[-1], frames[0])
        check_grads()
        key = r.last_ticket.hint_key
        rasterizer._PAIR_HINTS[key] = 16                        # stale hint: the next frame overflows its capacity
        with pytest.raises(PairCapacityExceeded):
            loss_fn().backward()
        assert all(leaves[k].grad is None for k in names)       # nothing reached the leaves
        assert rasterizer._PAIR_HINTS[key] >= r.last_ticket.pairs
        rasterizer._PAIR_HINTS[key] = 16
        n_before = len(frames)
        backward_or_retry(loss_fn)
        assert len(frames) == n_before + 2                      # rendered twice: overflow, then complete
        assert torch.equal(frames[-1], frames[0])
        check_grads()
        with torch.no_grad():                                   # no_grad calls stay on the blocking path
            c, _ = r(leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
                     rotations=leaves["rotations"])
        assert torch.equal(c, frames[0])
    finally:
        _cabi.set_option("bin_shift", -1)
        rasterizer.ADAPT_BIN_SIZE = True


def test_bin_size_policy_follows_the_splat_extent_and_never_changes_results():
    """The Python layer asks for bins of about three splat extents (per-call flags bits 8..11): small splats
    -> finer bins than the image-only default, large splats -> the default or coarser; images, radii and
    gradients are the same either way (the bin size is a pure performance knob)."""
    from robosimgs_b200 import rasterizer
    from robosimgs_b200.scenes import Scene
    W, H = 640, 480                                         # image-only default: 64-px bins (shift 2)
    assert rasterizer._default_bin_shift(H, W) == 2
    w = torch.rand(3, H, W, generator=torch.Generator().manual_seed(31))
    key = (0, 5000, H, W)
    picked = {}
    for label, mul in (("small", 0.2), ("large", 3.0)):
        sc, cam, rs = small_scene(P=5000, degree=1, W=W, H=H, big=0)
        sc = Scene(sc.means3D, sc.shs, sc.opacities, sc.scales * mul, sc.rotations, 1)
        results = []
        for adapt in (False, True):
            rasterizer.ADAPT_BIN_SIZE = adapt
            rasterizer._BIN_POLICY.pop(key, None)
            rasterizer._PAIR_HINTS.pop(key, None)
            try:
                gpu_render(sc, cam, 1, bg=(0.2, 0.1, 0.4))                      # first frame: the policy looks at it
                results.append(gpu_render(sc, cam, 1, bg=(0.2, 0.1, 0.4), grad_weight=w))
            finally:
                rasterizer.ADAPT_BIN_SIZE = True
        picked[label] = rasterizer._BIN_POLICY[key]["shift"]
        (c0, r0, g0), (c1, r1, g1) = results
        assert np.array_equal(c0, c1) and np.array_equal(r0, r1)
        for k in ("means3D", "shs", "opacities", "scales", "rotations"):
            assert max_rel_err(g1[k], g0[k]) < 1e-5, (label, k)
    assert picked["small"] == 1                             # finer than the default
    assert picked["large"] in (-1, 2, 3, 4)                 # default kept (or coarser)
    rasterizer._BIN_POLICY.pop(key, None)


def test_export_rgb8_matches_numpy():
    from robosimgs_b200 import export_rgb8
    g = torch.Generator().manual_seed(5)
    for H, W in ((48, 64), (23, 37)):
        col = (torch.rand(3, H, W, generator=g) * 1.4 - 0.2).cuda()
        out = export_rgb8(col).cpu().numpy()
        ref = np.rint(np.clip(col.cpu().numpy(), 0, 1) * 255.0).astype(np.uint8).transpose(1, 2, 0)
        assert out.shape == (H, W, 3) and np.abs(out.astype(int) - ref.astype(int)).max() <= 1
        assert (out != ref).mean() < 1e-3       # only exact .5 ties may differ


def test_bin_shift_gather_and_sort_modes_never_change_results():
    """Tuning knobs: every bin size (16..512 px), both record-gather modes and both pair-sort
    implementations (cooperative single-launch radix sort / CUB) give bit-identical images and
    radii, and the same gradients up to atomic-order noise."""
    from robosimgs_b200 import _cabi
    sc, cam, rs = small_scene(P=5000, degree=2, W=400, H=300)
    w = torch.rand(3, 300, 400, generator=torch.Generator().manual_seed(17))
    base = None
    try:
        # (binning, gather, sort, key width): bucketed per-bin sort / global radix sort (cooperative; CUB with
        # 32-bit quantised keys + exact tie repair -- used when there are <= 255 bins -- or 64-bit keys)
        for binning, gather, sort, keys in ((1, 1, 1, 32), (0, 1, 2, 32), (0, 0, 2, 32), (0, 1, 0, 32), (0, 1, 0, 64)):
            for shift in (-1, 0, 1, 2, 3, 5):
                _cabi.set_option("binning", binning)
                _cabi.set_option("gather", gather)
                _cabi.set_option("sort", sort)
                _cabi.set_option("sort_keys", keys)
                _cabi.set_option("bin_shift", shift)
                color, radii, grads = gpu_render(sc, cam, 2, bg=(0.2, 0.1, 0.4), grad_weight=w)
                if base is None:
                    base = (color, radii, grads)
                    continue
                assert np.array_equal(color, base[0]) and np.array_equal(radii, base[1]), (binning, gather, sort, keys, shift)
                for k in ("means3D", "shs", "opacities", "scales", "rotations"):
                    assert max_rel_err(grads[k], base[2][k]) < 1e-5, (binning, gather, sort, keys, shift, k)
    finally:
        _cabi.set_option("binning", -1)
        _cabi.set_option("gather", 1)
        _cabi.set_option("sort", 1)
        _cabi.set_option("sort_keys", 32)
        _cabi.set_option("bin_shift", -1)


def test_equal_depths_keep_index_order_in_both_binning_pipelines():
    """Depth ties must be composited in index order (a stable sort of index-ordered pairs).  A
    fronto-parallel plane gives thousands of bit-identical view depths per bin (long runs: the bucketed
    sort re-sorts the bin on the full key); a second scene quantises depth to a few values (short and
    medium runs: insertion).  Both pipelines must agree bit for bit and match the oracle."""
    from robosimgs_b200 import _cabi
    from robosimgs_b200.cameras import camera_look_at
    from robosimgs_b200.scenes import Scene, settings_from_camera
    g = torch.Generator().manual_seed(41)
    P = 6000
    cam = camera_look_at((0.0, 0.0, 3.0), (0.0, 0.0, 0.0), (0, 1, 0), 50.0, 160, 120)
    rs = settings_from_camera(cam, 0, bg=(0.1, 0.1, 0.1))
    xy = torch.rand(P, 2, generator=g) * 2.4 - 1.2
    rgb = torch.rand(P, 1, 3, generator=g)
    common = dict(shs=(rgb - 0.5) / 0.28209479177387814, opacities=torch.rand(P, 1, generator=g) * 0.5 + 0.2,
                  scales=torch.rand(P, 3, generator=g) * 0.05 + 0.02,
                  rotations=torch.tensor([[1.0, 0, 0, 0]]).repeat(P, 1))
    try:
        for levels in (1, 700):
            z = torch.zeros(P, 1) if levels == 1 else torch.randint(0, levels, (P, 1), generator=g).float() / 1024.0
            sc = Scene(torch.cat([xy, z], 1), common["shs"], common["opacities"], common["scales"], common["rotations"], 0)
            st = _oracle(rs, sc)
            assert len(np.unique(st.depths[st.radii > 0])) <= levels
            imgs = []
            # bucketed; global CUB sort with 32-bit keys (tie runs repaired: long runs by k_fix_long_runs, short
            # ones by insertion) and with 64-bit keys; cooperative sort
            for binning, sort, keys in ((1, 1, 32), (0, 0, 32), (0, 0, 64), (0, 2, 32)):
                _cabi.set_option("binning", binning)
                _cabi.set_option("sort", sort)
                _cabi.set_option("sort_keys", keys)
                color, radii, _ = gpu_render(sc, cam, 0, bg=(0.1, 0.1, 0.1))
                assert psnr(color, st.color) >= PSNR_MIN, (levels, binning, sort, keys)
                imgs.append(color)
            assert all(np.array_equal(imgs[0], im) for im in imgs[1:]), levels
    finally:
        _cabi.set_option("binning", -1)
        _cabi.set_option("sort", 1)
        _cabi.set_option("sort_keys", 32)


def test_one_and_four_pixel_compositing_kernels_agree():
    """render.cu (one pixel per thread), render4.cu (four pixels per thread; chosen automatically on large images) implement the
    same compositing rules: both meet the oracle bars on every bin size, each is bit-reproducible
    across bin sizes, and they agree with each other far inside the tolerance."""
    from oracle import gs_oracle
    from robosimgs_b200 import _cabi
    w = torch.rand(3, 150, 203, generator=torch.Generator().manual_seed(23))
    names = ("means3D", "shs", "opacities", "scales", "rotations")
    try:
        for P, op_scale in ((4000, 1.0), (6000, 0.1)):       # opaque early-out / long lists
            sc, cam, rs = small_scene(P=P, degree=3, W=203, H=150)      # ragged: W % 4 != 0 -> scalar pixel I/O
            sc.opacities.mul_(op_scale)
            st = _oracle(rs, sc)
            ref = gs_oracle.backward(st, w.numpy())
            per_mode = {}
            for mode in (1, 0):          # four pixels per thread, one pixel per thread
                _cabi.set_option("render", mode)
                base = None
                for shift in (-1, 0, 2):
                    _cabi.set_option("bin_shift", shift)
                    color, radii, grads = gpu_render(sc, cam, 3, bg=(0.2, 0.1, 0.4), grad_weight=w)
                    assert psnr(color, st.color) >= PSNR_MIN, (mode, shift)
                    _check_grads(grads, ref, names)
                    if base is None:
                        base = (color, radii)
                    assert np.array_equal(color, base[0]) and np.array_equal(radii, base[1]), (mode, shift)
                per_mode[mode] = (color, grads)
            assert psnr(per_mode[0][0], per_mode[1][0]) > 100.0
            for k in names:
                assert max_rel_err(per_mode[1][1][k], per_mode[0][1][k]) < 2e-4, k
        # W % 4 == 0: 128-bit pixel path of the four-pixel kernels, image not a multiple of the tile
        sc, cam, rs = small_scene(P=3000, degree=1, W=200, H=100)
        st = _oracle(rs, sc)
        w2 = torch.rand(3, 100, 200, generator=torch.Generator().manual_seed(29))
        ref = gs_oracle.backward(st, w2.numpy())
        _cabi.set_option("render", 1)
        _cabi.set_option("bin_shift", -1)
        color, radii, grads = gpu_render(sc, cam, 1, bg=(0.2, 0.1, 0.4), grad_weight=w2)
        assert psnr(color, st.color) >= PSNR_MIN
        _check_grads(grads, ref, names)
    finally:
        _cabi.set_option("render", -1)
        _cabi.set_option("bin_shift", -1)


def test_gsplat_style_rasterization_shim_matches_oracle():
    """gsplat signature (viewmats + Ks, near_plane 0.01, un-normalised quats, off-centre principal
    point, alphas) mapped onto the same kernels; checked against the oracle run with the equivalent
    inria-style settings."""
    from oracle import gs_oracle
    from robosimgs_b200.gsplat_compat import rasterization
    from robosimgs_b200.rasterizer import GaussianRasterizationSettings
    sc, cam, _ = small_scene(P=2500, degree=2, W=176, H=128, eye=(0.2, 0.1, 0.35), fov=70.0)   # splats inside 0.2
    W, H = 176, 128
    V = cam.viewmatrix.T.contiguous()                      # world->camera
    fx = W / (2 * cam.tanfovx); fy = H / (2 * cam.tanfovy)
    K = torch.tensor([[fx, 0, W / 2 + 7.25], [0, fy, H / 2 - 4.5], [0, 0, 1.0]])
    scale_q = torch.rand(2500, 1, generator=torch.Generator().manual_seed(2)) + 0.5
    dev = torch.device("cuda:0")
    leaf = lambda t: t.to(dev).clone().requires_grad_(True)
    means, quats, scales, opac, shs = leaf(sc.means3D), leaf(sc.rotations * scale_q), leaf(sc.scales), \
        leaf(sc.opacities.reshape(-1)), leaf(sc.shs)
    bgs = torch.tensor([[0.2, 0.1, 0.4]], device=dev)
    colors, alphas, meta = rasterization(means, quats, scales, opac, shs, V[None].to(dev), K[None].to(dev), W, H,
                                         sh_degree=2, backgrounds=bgs)
    assert colors.shape == (1, H, W, 3) and alphas.shape == (1, H, W, 1) and meta["radii"].shape == (1, 2500)
    assert meta["means2d"].shape == (1, 2500, 2)
    meta["means2d"].retain_grad()                          # what splatfacto's densification strategy does
    w = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(3))
    (colors[0] * w.to(dev)).sum().backward()
    # oracle with the equivalent settings
    from robosimgs_b200.gsplat_compat import _camera_matrices
    view_t, proj_t, campos, tfx, tfy = _camera_matrices(V, K, (float(fx), float(fy)), W, H, 0.01, 1000.0)
    rs = GaussianRasterizationSettings(H, W, tfx, tfy, torch.tensor([0.2, 0.1, 0.4]), 1.0, view_t, proj_t, 2, campos,
                                       False, False)
    qn = (sc.rotations * scale_q) / (sc.rotations * scale_q).norm(dim=1, keepdim=True)
    st = gs_oracle.forward(rs, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=qn,
                           dtype=np.float64, near_plane=0.01)
    assert ((st.depths > 0.01) & (st.depths < 0.2) & (st.radii > 0)).sum() > 10     # near-plane change matters
    assert psnr(colors[0].detach().cpu().numpy().transpose(2, 0, 1), st.color) >= PSNR_MIN
    assert np.abs(alphas[0, ..., 0].cpu().numpy() - (1 - st.final_T)).max() < 2e-3
    ref = gs_oracle.backward(st, w.numpy().transpose(2, 0, 1))
    assert max_rel_err(means.grad.cpu().numpy(), ref.means3D) < GRAD_TOL
    assert max_rel_err(shs.grad.cpu().numpy(), ref.shs) < GRAD_TOL
    assert max_rel_err(scales.grad.cpu().numpy(), ref.scales) < GRAD_TOL
    # meta["means2d"].grad is in PIXEL units (gsplat); the oracle reports it NDC-scaled (x 0.5 W, x 0.5 H)
    g2d = meta["means2d"].grad[0].cpu().numpy() * np.array([0.5 * W, 0.5 * H])
    assert max_rel_err(g2d, ref.means2D[:, :2]) < GRAD_TOL
    with pytest.raises(NotImplementedError):
        rasterization(means, quats, scales, opac, shs, V[None].to(dev), K[None].to(dev), W, H, sh_degree=2,
                      render_mode="RGB+ED")


def test_flat_input_layouts_and_forward_only_deferred_guard():
    """(a) opacities of shape [P] and shs of shape [P, M*3] are accepted by forward, so backward
    must hand back gradients of exactly those shapes; (b) forward_deferred never returns a frame that carries a
    grad_fn, even when called with grad enabled on leaves that require grad."""
    from oracle import gs_oracle
    from robosimgs_b200 import GaussianRasterizer
    from robosimgs_b200.scenes import settings_from_camera
    sc, cam, rs = small_scene(P=1200, degree=1, W=120, H=88)
    dev = torch.device("cuda:0")
    rs_dev = settings_from_camera(cam, 1, bg=(0.2, 0.1, 0.4), device=dev)
    leaf = lambda t: t.to(dev).clone().requires_grad_(True)
    means, opac, shs = leaf(sc.means3D), leaf(sc.opacities.reshape(-1)), leaf(sc.shs.reshape(sc.P, -1))
    scales, rots = leaf(sc.scales), leaf(sc.rotations)
    m2d = torch.zeros_like(means, requires_grad=True)
    w = torch.rand(3, 88, 120, generator=torch.Generator().manual_seed(21))
    color, _ = GaussianRasterizer(rs_dev)(means, m2d, opac, shs=shs, scales=scales, rotations=rots)
    (color * w.to(dev)).sum().backward()
    assert opac.grad.shape == opac.shape and shs.grad.shape == shs.shape
    ref = gs_oracle.backward(_oracle(rs, sc), w.numpy())
    assert max_rel_err(opac.grad.cpu().numpy(), ref.opacities.reshape(-1)) < GRAD_TOL
    assert max_rel_err(shs.grad.cpu().numpy().reshape(ref.shs.shape), ref.shs) < GRAD_TOL
    frame, _, ticket = GaussianRasterizer(rs_dev).forward_deferred(means, m2d, opac, shs=shs, scales=scales,
                                                                   rotations=rots)
    assert frame.grad_fn is None and not frame.requires_grad
    assert ticket.ok()
    assert torch.equal(frame, color.detach())


@pytest.mark.parametrize("P,spread", [(3000, 0.0005), (40000, 0.0005), (160000, 0.0), (60000, 0.6)])
def test_bucket_sort_size_classes_match_the_global_sort(P, spread):
    """Depth-sliced bucket binning (bucket.cu) against the library-sort pipeline, bit for bit, on walls that face the
    camera -- nearly all pairs of a bin fall into one or two depth slices, so the buckets run through every size
    class: the in-register warp sort (<= 512 keys), the one-CTA radix passes with keys in registers (<= 4096,
    <= 12288) and the any-length passes; spread = 0 makes every depth bit-identical (pure index order), a wide
    spread exercises many small buckets.  Both bin sizes (16 px: many bins, few slices; 128 px: the reverse)."""
    from robosimgs_b200 import _cabi
    from robosimgs_b200.cameras import camera_look_at
    from robosimgs_b200.scenes import Scene
    g = torch.Generator().manual_seed(P)
    cam = camera_look_at((0.0, 0.0, 3.0), (0.0, 0.0, 0.0), (0, 1, 0), 50.0, 320, 256)
    xy = torch.rand(P, 2, generator=g) * 3.0 - 1.5
    z = (torch.rand(P, 1, generator=g) - 0.5) * spread
    rgb = torch.rand(P, 1, 3, generator=g)
    sc = Scene(torch.cat([xy, z], 1), (rgb - 0.5) / 0.28209479177387814, torch.rand(P, 1, generator=g) * 0.3 + 0.05,
               torch.rand(P, 3, generator=g) * 0.03 + 0.01, torch.tensor([[1.0, 0, 0, 0]]).repeat(P, 1), 0)
    try:
        for shift in (0, 3):
            _cabi.set_option("bin_shift", shift)
            out = []
            for binning in (1, 0):
                _cabi.set_option("binning", binning)
                _cabi.set_option("sort", 0)
                _cabi.set_option("sort_keys", 64)
                color, radii, _ = gpu_render(sc, cam, 0, bg=(0.1, 0.1, 0.1))
                out.append(color)
            assert np.array_equal(out[0], out[1]), (P, spread, shift)
    finally:
        _cabi.set_option("binning", -1)
        _cabi.set_option("bin_shift", -1)
        _cabi.set_option("sort", 1)
        _cabi.set_option("sort_keys", 32)


def test_record_slab_tma_ring_is_bit_identical():
    """Option "slab": the four-pixel compositing kernels fill their ring with TMA bulk copies from the bin-ordered record
    slab (k_build_slab) instead of per-record LDGSTS gathers -- images, radii and gradients must not change; long lists
    (many chunks per tile, early exits with copies in flight) and tiny ones."""
    from oracle import gs_oracle
    from robosimgs_b200 import _cabi
    sc, cam, rs = small_scene(P=6000, degree=1, W=320, H=208, opacity_boost=2.0)
    sc.scales[:300] *= 6
    w = torch.rand(3, 208, 320, generator=torch.Generator().manual_seed(31))
    res = {}
    try:
        _cabi.set_option("render", 1)
        for slab in (0, 1):
            _cabi.set_option("slab", slab)
            for shift in (0, 2):
                _cabi.set_option("bin_shift", shift)
                res[(slab, shift)] = gpu_render(sc, cam, 1, bg=(0.2, 0.1, 0.4), grad_weight=w)
        base = res[(0, 0)]
        for key, (color, radii, grads) in res.items():
            assert np.array_equal(color, base[0]) and np.array_equal(radii, base[1]), key
            for k in grads:
                assert max_rel_err(grads[k], base[2][k]) < 1e-5, (key, k)
        ref = gs_oracle.backward(_oracle(rs, sc), w.numpy())
        _check_grads(res[(1, 2)][2], ref, ("means3D", "shs", "opacities", "scales", "rotations", "means2D"))
    finally:
        _cabi.set_option("slab", 0)
        _cabi.set_option("bin_shift", -1)
        _cabi.set_option("render", -1)


@pytest.mark.parametrize("degree,precomp", [(3, False), (1, False), (0, True)])
def test_dense_warp_projection_kernels_are_bit_identical(degree, precomp):
    """Option "project": the projection kernels with warp-level stream compaction (cull -> geometry -> colour, dense
    warps) against the one-thread-per-Gaussian kernels -- same arithmetic per Gaussian, so images and gradients agree to
    the last bit or two and radii exactly; a scene where most Gaussians are culled (behind the camera / off screen), ragged
    chunk ends (P not a multiple of 128), both binning pipelines."""
    from robosimgs_b200 import _cabi
    sc, cam, rs = small_scene(P=5003, degree=degree, W=200, H=136, eye=(0.2, 0.1, 0.6), fov=65.0)
    w = torch.rand(3, 136, 200, generator=torch.Generator().manual_seed(17))
    kw = {}
    if precomp:
        kw["colors_precomp"] = torch.rand(5003, 3, generator=torch.Generator().manual_seed(18))
    out = {}
    try:
        for binning in (1, 0):
            _cabi.set_option("binning", binning)
            for mode in (0, 1):
                _cabi.set_option("project", mode)
                out[(binning, mode)] = gpu_render(sc, cam, degree, bg=(0.2, 0.1, 0.4), grad_weight=w, **kw)
        base = out[(1, 0)]
        assert 0.05 < (base[1] > 0).mean() < 0.9
        for key, (color, radii, grads) in out.items():
            # the two binning pipelines feed identical lists to the same kernels: bit for bit; the two projection
            # kernels are separate compilations of the same expressions (the compiler contracts a*b+c per kernel), so
            # their records may differ in the last bit: colours to 1e-6, radii exactly
            if key[1] == 0:
                assert np.array_equal(color, base[0]), key
            assert np.abs(color - base[0]).max() < 1e-6 and np.array_equal(radii, base[1]), key
            for k in grads:
                assert np.array_equal(grads[k], base[2][k]) or max_rel_err(grads[k], base[2][k]) < 1e-5, (key, k)      # (the adjoint's RED order varies)
    finally:
        _cabi.set_option("project", 0)
        _cabi.set_option("binning", -1)


@pytest.mark.parametrize("W,H,shift,expect_bucketed", [(512, 512, 0, True), (1104, 1104, 0, True), (1920, 1080, 0, True),
                                                       (1920, 1080, 2, True), (4096, 4096, 0, False)])
def test_bin_count_limits_of_the_bucketed_binning(W, H, shift, expect_bucketed):
    """The bucket scan runs one CTA per bin with a decoupled look-back (aggregate / inclusive-prefix words) over the
    lower bins; it serves up to 65534 bins (bin ids travel in 16 bits), images with more bins take the library-sort
    pipeline.  Few bins with thousands of slices, thousands of bins with a handful of slices each (8160 scan CTAs at
    1080p with 16-px bins), and the far side of the limit: against the oracle, and bit for bit against 128-px bins."""
    from robosimgs_b200 import _cabi
    from robosimgs_b200.cameras import camera_look_at
    from robosimgs_b200.scenes import cube_scene, settings_from_camera
    bins = -(-((W + 15) // 16) // (1 << shift)) * -(-((H + 15) // 16) // (1 << shift))
    assert (bins < 65535 and bins * 4 <= 512 * 1024) == expect_bucketed
    sc, _ = cube_scene(P=20000, seed=9, degree=1)
    cam = camera_look_at((0.4, 0.3, 3.2), (0, 0, 0), (0, 1, 0), 55.0, W, H)
    rs = settings_from_camera(cam, 1, bg=(0.1, 0.2, 0.3))
    try:
        _cabi.set_option("bin_shift", shift)
        color, radii, _ = gpu_render(sc, cam, 1, bg=(0.1, 0.2, 0.3))
        _cabi.set_option("bin_shift", 3)
        color3, _, _ = gpu_render(sc, cam, 1, bg=(0.1, 0.2, 0.3))
    finally:
        _cabi.set_option("bin_shift", -1)
    assert np.array_equal(color, color3)
    st = _oracle(rs, sc, np.float32)
    assert psnr(color, st.color) >= PSNR_MIN
    assert (radii != st.radii).mean() <= 1e-3


def test_backward_zero_fill_overlap_gives_the_same_gradients():
    """Option "bwd_overlap": gradient tensors zero-filled on a side stream under the compositing adjoint, projection
    adjoint writes only the rows of visible Gaussians -- every gradient element must equal the one-pass result
    (culled Gaussians: exact zeros), for SH and precomputed-colour inputs."""
    from robosimgs_b200 import _cabi
    sc, cam, rs = small_scene(P=5003, degree=2, W=200, H=136, eye=(0.2, 0.1, 0.6), fov=65.0)
    w = torch.rand(3, 136, 200, generator=torch.Generator().manual_seed(23))
    cols = torch.rand(5003, 3, generator=torch.Generator().manual_seed(24))
    try:
        for kw in ({}, {"colors_precomp": cols}):
            out = {}
            for mode in (0, 1, 1):
                _cabi.set_option("bwd_overlap", mode)
                out[mode] = gpu_render(sc, cam, 2, bg=(0.2, 0.1, 0.4), grad_weight=w, **kw)
            assert np.array_equal(out[0][0], out[1][0])
            culled = out[0][1] <= 0
            assert culled.mean() > 0.2
            for k in out[0][2]:
                a, b = out[0][2][k], out[1][2][k]
                assert max_rel_err(b, a) < 1e-5, k
                assert not np.any(b.reshape(5003, -1)[culled]), k          # exact zeros where nothing was rendered
    finally:
        _cabi.set_option("bwd_overlap", 0)
