"""Measured comparator for BASELINE.json's ">= 1.5x the reference rasterizer" target:
baseline/naive_simt.cu -- a straight SIMT restatement of the published 3DGS rasterizer design, NOT product
code -- against libb200gs on the same GPU, same C3 scene and camera.  Checks first that both render the same
image.  Writes gpurun_out/naive_comparator.json.   python tools/compare_naive.py [P]"""
import ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from robosimgs_b200 import GaussianRasterizer, _cabi
from robosimgs_b200.losses import mse_loss
from robosimgs_b200.scenes import room_scene, room_target, settings_from_camera

SO = os.path.join(ROOT, "baseline", "libnaive3dgs.so")
if not os.path.exists(SO):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "-Xcompiler", "-fPIC", "-shared", "-o", SO, os.path.join(ROOT, "baseline", "naive_simt.cu"), "-lcudart"])
N = C.CDLL(SO)
N.naive_preprocess.restype = C.c_int64
N.naive_temp_bytes.restype = C.c_size_t
N.naive_temp_bytes.argtypes = [C.c_int, C.c_int64]

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
sc, cam = room_scene(P)
rs = settings_from_camera(cam, 3, device=dev)
H, W = cam.image_height, cam.image_width
t = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
target = room_target(W, H).to(dev)
p = lambda x: C.c_void_p(x.data_ptr())
f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
u32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
gx, gy = (W + 15) // 16, (H + 15) // 16

radii, xy, depths, co, rgb = u32(P), f32(P, 2), f32(P), f32(P, 4), f32(P, 3)
tiles, offsets = u32(P), u32(P)
finalT, ncontrib, out = f32(H, W), u32(H, W), f32(3, H, W)
ranges = u32(gx * gy, 2)
bufs = {}

def naive_forward(ev=None):
    tb = N.naive_temp_bytes(P, bufs.get("D", 1))
    if bufs.get("temp") is None or bufs["temp"].numel() < tb:
        bufs["temp"] = torch.empty(tb, dtype=torch.uint8, device=dev)
    if ev: ev[0].record()
    D = N.naive_preprocess(C.c_int(P), C.c_int(3), C.c_int(16), C.c_int(H), C.c_int(W), C.c_float(cam.tanfovx),
                           C.c_float(cam.tanfovy), p(rs.viewmatrix), p(rs.projmatrix), p(rs.campos), p(rs.bg),
                           p(t["means3D"]), p(t["scales"]), p(t["rotations"]), p(t["opacities"]), p(t["shs"]),
                           p(radii), p(xy), p(depths), p(co), p(rgb), p(tiles), p(offsets), p(bufs["temp"]),
                           C.c_size_t(bufs["temp"].numel()), stream)
    assert D >= 0
    if ev: ev[1].record()
    if bufs.get("D", -1) < D:                      # grow the binning buffers like the resize callbacks would
        bufs["D"] = int(D * 1.05)
        bufs["keys"] = torch.empty(bufs["D"], dtype=torch.int64, device=dev)
        bufs["keys_s"] = torch.empty(bufs["D"], dtype=torch.int64, device=dev)
        bufs["vals"], bufs["vals_s"] = u32(bufs["D"]), u32(bufs["D"])
        tb = N.naive_temp_bytes(P, bufs["D"])
        bufs["temp"] = torch.empty(tb, dtype=torch.uint8, device=dev)
    rc = N.naive_bin_and_render(C.c_int(P), C.c_int(H), C.c_int(W), C.c_int64(D), p(rs.bg), p(radii), p(xy), p(depths),
                                p(co), p(rgb), p(offsets), p(bufs["keys"]), p(bufs["keys_s"]), p(bufs["vals"]),
                                p(bufs["vals_s"]), p(ranges), p(bufs["temp"]), C.c_size_t(bufs["temp"].numel()),
                                p(finalT), p(ncontrib), p(out), stream)
    assert rc == 0
    if ev: ev[2].record()
    return D

acc = f32(P, 12)
def naive_render_bwd(g):
    rc = N.naive_render_backward(C.c_int(H), C.c_int(W), p(rs.bg), p(xy), p(co), p(rgb), p(bufs["vals_s"]), p(ranges),
                                 p(finalT), p(ncontrib), p(g), p(acc), C.c_int(P), stream)
    assert rc == 0

def time_ms(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

# ---- same image? ----
r = GaussianRasterizer(rs)
m2d = torch.zeros_like(t["means3D"])
with torch.no_grad():
    ours, _ = r(t["means3D"], m2d, t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
D_naive = naive_forward()
torch.cuda.synchronize()
mse = float(((ours - out).double() ** 2).mean())
psnr = 10 * torch.log10(torch.tensor(1.0 / max(mse, 1e-30))).item()
assert psnr > 60.0, psnr

# ---- timings ----
res = {"P": P, "image": [W, H], "D_naive": int(D_naive), "psnr_naive_vs_b200gs_dB": psnr}
res["naive_fwd_ms"] = time_ms(naive_forward, 20)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
naive_forward(ev); torch.cuda.synchronize()
res["naive_stage_ms"] = {"preprocess+scan(+host sync)": ev[0].elapsed_time(ev[1]), "duplicate+sort+ranges+render": ev[1].elapsed_time(ev[2])}
g = (2.0 / out.numel()) * (out - target)
res["naive_render_bwd_ms"] = time_ms(lambda: naive_render_bwd(g), 10)
with torch.no_grad():
    res["b200gs_fwd_ms"] = time_ms(lambda: r(t["means3D"], m2d, t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"]), 30)
leaves = {k: v.clone().requires_grad_(True) for k, v in t.items()}
m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
def train():
    for v in leaves.values(): v.grad = None
    c, _ = r(leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    mse_loss(c, target).backward()
res["b200gs_train_ms"] = time_ms(train, 15)
_cabi.profile_enable(True); _cabi.profile_read(True)
train(); torch.cuda.synchronize()
st = _cabi.profile_read(True); _cabi.profile_enable(False)
pb = st["project_bwd"][0] / max(st["project_bwd"][1], 1)
res["b200gs_project_bwd_ms"] = pb
res["naive_train_ms_estimate"] = res["naive_fwd_ms"] + res["naive_render_bwd_ms"] + pb
res["speedup_fwd"] = res["naive_fwd_ms"] / res["b200gs_fwd_ms"]
res["speedup_train"] = res["naive_train_ms_estimate"] / res["b200gs_train_ms"]
res["note"] = ("naive = straight SIMT restatement of the published design (baseline/naive_simt.cu), single stream; its train "
               "estimate = its forward + its per-pixel-atomic compositing adjoint + libb200gs's projection adjoint (shared)")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/naive_comparator.json", "w"), indent=1)
print(json.dumps(res))
