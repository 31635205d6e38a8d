"""Driver for baseline/naive_simt.cu (libnaive3dgs.so): the plain-SIMT restatement of the published 3DGS rasterizer
that stands in for the real third-party library as the GPU yardstick of BASELINE.json's ">= 1.5x the reference
rasterizer" target.  NOT product code: only bench.py's `gpu_comparator` leg and tools/compare_naive.py import it."""
import ctypes as C
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnaive3dgs.so")


class NaiveRasterizer:
    """Forward of the naive restatement for one fixed scene / image size (buffers grown like the public
    binding's resize callbacks)."""

    def __init__(self, tens: dict, height: int, width: int, sh_degree: int = 3):
        if not os.path.exists(_SO):
            raise FileNotFoundError(f"{_SO}: run __graft_entry__.build()")
        N = self.N = C.CDLL(_SO)
        N.naive_preprocess.restype = C.c_int64
        N.naive_temp_bytes.restype = C.c_size_t
        N.naive_temp_bytes.argtypes = [C.c_int, C.c_int64]
        self.t, self.H, self.W, self.deg = tens, int(height), int(width), int(sh_degree)
        self.dev = dev = tens["means3D"].device
        P = self.P = tens["means3D"].shape[0]
        self.M = tens["shs"].shape[1]
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        u32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
        self.u32 = u32
        gx, gy = (self.W + 15) // 16, (self.H + 15) // 16
        self.radii, self.xy, self.depths, self.co, self.rgb = u32(P), f32(P, 2), f32(P), f32(P, 4), f32(P, 3)
        self.tiles, self.offsets = u32(P), u32(P)
        self.finalT, self.ncontrib, self.out = f32(self.H, self.W), u32(self.H, self.W), f32(3, self.H, self.W)
        self.ranges = u32(gx * gy, 2)
        self.bufs = {}

    def forward(self, rs, events=None):
        """rs: GaussianRasterizationSettings with device tensors.  Returns the pair count D; image in self.out."""
        N, b, t, P, H, W = self.N, self.bufs, self.t, self.P, self.H, self.W
        p = lambda x: C.c_void_p(x.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        tb = N.naive_temp_bytes(P, b.get("D", 1))
        if b.get("temp") is None or b["temp"].numel() < tb:
            b["temp"] = torch.empty(tb, dtype=torch.uint8, device=self.dev)
        if events:
            events[0].record()
        D = N.naive_preprocess(C.c_int(P), C.c_int(self.deg), C.c_int(self.M), C.c_int(H), C.c_int(W),
                               C.c_float(rs.tanfovx), C.c_float(rs.tanfovy), p(rs.viewmatrix), p(rs.projmatrix),
                               p(rs.campos), p(rs.bg), p(t["means3D"]), p(t["scales"]), p(t["rotations"]),
                               p(t["opacities"]), p(t["shs"]), p(self.radii), p(self.xy), p(self.depths), p(self.co),
                               p(self.rgb), p(self.tiles), p(self.offsets), p(b["temp"]), C.c_size_t(b["temp"].numel()),
                               stream)
        assert D >= 0
        if events:
            events[1].record()
        if b.get("D", -1) < D:
            b["D"] = int(D * 1.05)
            b["keys"] = torch.empty(b["D"], dtype=torch.int64, device=self.dev)
            b["keys_s"] = torch.empty(b["D"], dtype=torch.int64, device=self.dev)
            b["vals"], b["vals_s"] = self.u32(b["D"]), self.u32(b["D"])
            b["temp"] = torch.empty(N.naive_temp_bytes(P, b["D"]), dtype=torch.uint8, device=self.dev)
        rc = N.naive_bin_and_render(C.c_int(P), C.c_int(H), C.c_int(W), C.c_int64(D), p(rs.bg), p(self.radii), p(self.xy),
                                    p(self.depths), p(self.co), p(self.rgb), p(self.offsets), p(b["keys"]), p(b["keys_s"]),
                                    p(b["vals"]), p(b["vals_s"]), p(self.ranges), p(b["temp"]),
                                    C.c_size_t(b["temp"].numel()), p(self.finalT), p(self.ncontrib), p(self.out), stream)
        assert rc == 0
        if events:
            events[2].record()
        return int(D)

    def render_backward(self, rs, dL_dcolor, acc):
        p = lambda x: C.c_void_p(x.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        rc = self.N.naive_render_backward(C.c_int(self.H), C.c_int(self.W), p(rs.bg), p(self.xy), p(self.co), p(self.rgb),
                                          p(self.bufs["vals_s"]), p(self.ranges), p(self.finalT), p(self.ncontrib),
                                          p(dL_dcolor), p(acc), C.c_int(self.P), stream)
        assert rc == 0
