"""B200GSContext (include/b200gs.h): the no-stall protocol behind the C ABI, driven through ctypes exactly as a C host
would -- context-owned scratch arenas, pair-capacity hint, deferred tickets, adaptive bin size, backward -- against the
Python operator path (bit-identical images) and the fp64 oracle (gradients)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import max_rel_err, psnr, small_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(built):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


class Ctx:
    def __init__(self):
        from robosimgs_b200 import _cabi
        self.cabi, self.L = _cabi, _cabi.lib()
        self.h = C.c_void_p()
        _cabi.check(self.L.b200gs_context_create(C.byref(self.h)))

    def close(self):
        self.cabi.check(self.L.b200gs_context_destroy(self.h))

    def forward(self, sc, rs, degree, defer=False, flags=0):
        """sc: dict of CUDA tensors; rs: settings with CUDA tensors.  Context arenas for all scratch (NULL allocators)."""
        cabi, L = self.cabi, self.L
        P, M = sc["means3D"].shape[0], sc["shs"].shape[1]
        prm = cabi.B200GSParams(P, degree, M, rs.image_height, rs.image_width, rs.tanfovx, rs.tanfovy, rs.scale_modifier,
                                0, 0, 0.0, flags, 0)
        color = torch.empty((3, rs.image_height, rs.image_width), device="cuda")
        radii = torch.empty(P, dtype=torch.int32, device="cuda")
        p = lambda t: C.c_void_p(t.data_ptr())
        null = cabi.B200GSAlloc(None, cabi.RESIZE_FN(0))
        D, ticket, flags = C.c_int32(-1), C.c_int64(-7), C.c_int32(-1)
        cabi.check(L.b200gs_context_forward(self.h, C.byref(prm), p(rs.bg), p(rs.viewmatrix), p(rs.projmatrix), p(rs.campos),
                                            p(sc["means3D"]), p(sc["shs"]), None, p(sc["opacities"]), p(sc["scales"]),
                                            p(sc["rotations"]), None, p(color), p(radii), null, null, null,
                                            C.c_int32(1 if defer else 0), C.byref(D), C.byref(ticket), C.byref(flags),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return color, radii, D.value, ticket.value, flags.value, prm

    def wait(self, ticket):
        D, ok = C.c_int32(-1), C.c_int32(-1)
        self.cabi.check(self.L.b200gs_context_ticket_wait(self.h, C.c_int64(ticket), C.byref(D), C.byref(ok)))
        return D.value, bool(ok.value)

    def query(self, P, H, W):
        t, s = C.c_int64(-1), C.c_int32(-9)
        self.cabi.check(self.L.b200gs_context_query(self.h, P, H, W, C.byref(t), C.byref(s)))
        return t.value, s.value


def _operator(sc_cpu, cam, degree, bg):
    from helpers import gpu_render
    color, radii, _ = gpu_render(sc_cpu, cam, degree, bg=bg)
    return color, radii


def _dev_scene(sc):
    return {k: getattr(sc, k).cuda().contiguous() for k in ("means3D", "shs", "opacities", "scales", "rotations")}


def test_context_sync_then_deferred_frames_match_the_operator():
    from robosimgs_b200.cameras import camera_look_at
    from robosimgs_b200.scenes import settings_from_camera
    sc, cam, _ = small_scene(P=4000, degree=2, W=320, H=240)
    dsc = _dev_scene(sc)
    bg = (0.2, 0.1, 0.4)
    ref, ref_radii = _operator(sc, cam, 2, bg)
    rs = settings_from_camera(cam, 2, bg=bg, device="cuda")
    ctx = Ctx()
    try:
        color, radii, D, ticket, flags, _ = ctx.forward(dsc, rs, 2, defer=True)        # no hint yet: behaves synchronously
        torch.cuda.synchronize()
        assert ticket == -1 and D > 0
        assert np.array_equal(color.cpu().numpy(), ref) and np.array_equal(radii.cpu().numpy(), ref_radii)
        tracked, shift = ctx.query(4000, 240, 320)
        if shift != -1:                 # the policy picked another bin size: the tracked pair count starts over
            assert tracked == 0
            color, radii, D, ticket, flags, _ = ctx.forward(dsc, rs, 2, defer=True)
            torch.cuda.synchronize()
            assert ticket == -1 and np.array_equal(color.cpu().numpy(), ref)
            tracked, _ = ctx.query(4000, 240, 320)
        assert tracked == D
        seen = []
        for _ in range(3):                                                              # now the host never waits
            color, radii, _, ticket, flags2, _ = ctx.forward(dsc, rs, 2, defer=True)
            assert ticket >= 0
            seen.append((color, ticket))
        for color, ticket in seen:
            D2, ok = ctx.wait(ticket)
            assert ok and D2 == D
            assert np.array_equal(color.cpu().numpy(), ref)
        with pytest.raises(Exception):
            ctx.wait(seen[0][1])                                                        # a ticket is good for one wait
        # an abrupt view change: many more pairs than the hint -> the deferred frame reports incomplete, the exact
        # re-render is right and the hint recovers
        cam2 = camera_look_at((0.1, 0.05, 0.9), (0, 0, 0), (0, 1, 0), 60.0, 320, 240)
        rs2 = settings_from_camera(cam2, 2, bg=bg, device="cuda")
        ref2, _ = _operator(sc, cam2, 2, bg)
        _, _, _, ticket, _, _ = ctx.forward(dsc, rs2, 2, defer=True)
        D3, ok = ctx.wait(ticket)
        if not ok:
            color, _, D4, t4, _, _ = ctx.forward(dsc, rs2, 2, defer=False)
            assert t4 == -1 and D4 == D3
        else:
            color, _, _, _, _, _ = ctx.forward(dsc, rs2, 2, defer=False)
        torch.cuda.synchronize()
        assert np.array_equal(color.cpu().numpy(), ref2)
        assert ctx.query(4000, 240, 320)[0] >= D3
    finally:
        ctx.close()


def test_context_backward_matches_oracle_and_bin_policy_never_changes_results():
    from oracle import gs_oracle
    from robosimgs_b200.scenes import settings_from_camera
    sc, cam, rs_cpu = small_scene(P=5000, degree=1, W=640, H=400, big=0)
    sc.scales.mul_(0.35)                                   # small splats: the policy should ask for small bins
    dsc = _dev_scene(sc)
    rs = settings_from_camera(cam, 1, bg=(0.2, 0.1, 0.4), device="cuda")
    st = gs_oracle.forward(rs_cpu, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations,
                           dtype=np.float64)
    w = torch.rand(3, 400, 640, generator=torch.Generator().manual_seed(5))
    ref = gs_oracle.backward(st, w.numpy())
    ctx = Ctx()
    try:
        color0, radii, D, _, flags0, prm = ctx.forward(dsc, rs, 1)            # first frame: automatic bins, policy looks
        _, shift = ctx.query(5000, 400, 640)
        assert shift in (1, 2)                                               # 32- or 64-px bins for ~10-px splats
        color1, radii, D1, _, flags1, prm = ctx.forward(dsc, rs, 1)           # second frame: policy-chosen bins
        assert flags1 == ((shift + 1) << 8) and flags0 == 0
        torch.cuda.synchronize()
        assert torch.equal(color0, color1)
        assert psnr(color1.cpu().numpy(), st.color) >= 60.0
        # backward of the frame that lives in the context's arenas
        L, cabi = ctx.L, ctx.cabi
        p = lambda t: C.c_void_p(t.data_ptr())
        P, M = 5000, dsc["shs"].shape[1]
        g = {k: torch.empty_like(v) for k, v in dsc.items()}
        g2d = torch.empty(P, 3, device="cuda")
        dL = w.cuda().contiguous()
        cabi.check(L.b200gs_context_backward(ctx.h, C.byref(prm), C.c_int32(flags1), p(rs.bg), p(rs.viewmatrix), p(rs.projmatrix),
                                             p(rs.campos), p(dsc["means3D"]), p(dsc["shs"]), None, p(dsc["opacities"]),
                                             p(dsc["scales"]), p(dsc["rotations"]), None, p(radii), None, None, None,
                                             C.c_int32(D1), p(dL), p(g["means3D"]), p(g2d), p(g["shs"]), None, p(g["opacities"]),
                                             p(g["scales"]), p(g["rotations"]), None,
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        for k in ("means3D", "shs", "opacities", "scales", "rotations"):
            r = getattr(ref, k)
            assert max_rel_err(g[k].cpu().numpy().reshape(r.shape), r) < 1e-3, k
        assert max_rel_err(g2d.cpu().numpy(), ref.means2D) < 1e-3
        # training without a host wait, C side: deferred forward (adjoint state kept), backward launched at once with any
        # positive num_rendered (the adjoint only needs to know that there are pairs), THEN the ticket is validated
        g_def = {k: torch.empty_like(v) for k, v in dsc.items()}
        g2d_def = torch.empty(P, 3, device="cuda")
        color3, radii3, _, ticket3, flags3, prm3 = ctx.forward(dsc, rs, 1, defer=True)
        assert ticket3 >= 0 and flags3 == flags1
        cabi.check(L.b200gs_context_backward(ctx.h, C.byref(prm3), C.c_int32(flags3), p(rs.bg), p(rs.viewmatrix), p(rs.projmatrix),
                                             p(rs.campos), p(dsc["means3D"]), p(dsc["shs"]), None, p(dsc["opacities"]),
                                             p(dsc["scales"]), p(dsc["rotations"]), None, p(radii3), None, None, None,
                                             C.c_int32(1), p(dL), p(g_def["means3D"]), p(g2d_def), p(g_def["shs"]), None,
                                             p(g_def["opacities"]), p(g_def["scales"]), p(g_def["rotations"]), None,
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        D3, ok3 = ctx.wait(ticket3)
        assert ok3 and D3 == D1
        torch.cuda.synchronize()
        assert torch.equal(color3, color1)
        for k in ("means3D", "shs", "opacities", "scales", "rotations"):
            r = getattr(ref, k)
            assert max_rel_err(g_def[k].cpu().numpy().reshape(r.shape), r) < 1e-3, k
        # B200GS_FORWARD_ONLY travels through the context: same frame, same radii, and the adjoint refuses to run on it
        color2, radii2, D2, _, flags2, prm2 = ctx.forward(dsc, rs, 1, flags=cabi.FORWARD_ONLY)
        torch.cuda.synchronize()
        assert flags2 == (flags1 | cabi.FORWARD_ONLY) and D2 == D1
        assert torch.equal(color2, color1) and torch.equal(radii2, radii)
        # ... and so does B200GS_OUT_RGB8: the first H*W*3 bytes of the output buffer are the 8-bit frame
        from robosimgs_b200 import export_rgb8
        color4, _, _, _, flags4, _ = ctx.forward(dsc, rs, 1, flags=cabi.FORWARD_ONLY | cabi.OUT_RGB8)
        torch.cuda.synchronize()
        assert flags4 == (flags1 | cabi.FORWARD_ONLY | cabi.OUT_RGB8)
        H_, W_ = rs.image_height, rs.image_width
        frame8 = color4.view(torch.uint8).reshape(-1)[: H_ * W_ * 3].view(H_, W_, 3)
        assert torch.equal(frame8, export_rgb8(color1))
        with pytest.raises(cabi.B200GSError, match="FORWARD_ONLY"):
            cabi.check(L.b200gs_context_backward(ctx.h, C.byref(prm2), C.c_int32(flags2), p(rs.bg), p(rs.viewmatrix),
                                                 p(rs.projmatrix), p(rs.campos), p(dsc["means3D"]), p(dsc["shs"]), None,
                                                 p(dsc["opacities"]), p(dsc["scales"]), p(dsc["rotations"]), None, p(radii2),
                                                 None, None, None, C.c_int32(D2), p(dL), p(g["means3D"]), p(g2d), p(g["shs"]),
                                                 None, p(g["opacities"]), p(g["scales"]), p(g["rotations"]), None,
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    finally:
        ctx.close()


def test_policy_rules_are_the_ones_the_operator_layer_uses():
    from robosimgs_b200 import _cabi
    L = _cabi.lib()
    assert L.b200gs_policy_pair_capacity(0) == 0 and L.b200gs_policy_pair_capacity(1_000_000) == 1_000_000 + 62_500 + 32_768
    assert L.b200gs_policy_bin_shift(int(2.08 * 400), 400, 3, 0.9) == 3
    assert L.b200gs_policy_bin_shift(int(2.08 * 400), 400, 3, 0.9999) == 4
    assert L.b200gs_policy_bin_shift(int(1.16 * 400), 400, 3, -1.0) == 1
    assert L.b200gs_policy_bin_shift(0, 400, 3, 0.5) == 3
