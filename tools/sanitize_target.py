"""Small fwd+bwd cases for compute-sanitizer (memcheck / racecheck / synccheck): both compositing kernel
families, both record-gather modes and the TMA-fed slab ring, both binning pipelines (depth-sliced buckets: warp sort,
one-CTA radix for fat segments; library sort with 32- and 64-bit keys, tie repair incl. long runs; cooperative sort),
bin shifts auto/0/2, long lists crossing ring stages, ragged image, fused SSIM / Adam / photometric loss."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import small_scene, gpu_render
from robosimgs_b200 import _cabi
from robosimgs_b200.scenes import Scene
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
cases = ((3000, 200, 136, 3, -1), (6000, 64, 48, 0, 0), (500, 37, 23, 1, 2))
modes = [dict(render=1, gather=1, binning=0, sort=1, sort_keys=32, slab=0), dict(render=0, gather=1, binning=0, sort=0, sort_keys=32, slab=0),
         dict(render=0, gather=0, binning=0, sort=2, sort_keys=64, slab=0), dict(render=1, gather=1, binning=1, sort=1, sort_keys=32, slab=0),
         dict(render=1, gather=1, binning=0, sort=0, sort_keys=64, slab=0), dict(render=1, gather=1, binning=1, sort=1, sort_keys=32, slab=1),
         dict(render=0, gather=1, binning=1, sort=1, sort_keys=32, slab=0)]
if quick:
    modes, cases = [modes[0], modes[3], modes[5]], cases[:2]
for m in modes:
    for k, v in m.items():
        _cabi.set_option(k, v)
    for (P, W, H, deg, shift) in cases:
        _cabi.set_option("bin_shift", shift)
        sc, cam, rs = small_scene(P=P, degree=deg, W=W, H=H)
        if P == 6000:
            sc.opacities.mul_(0.15)
        w = torch.rand(3, H, W)
        color, radii, grads = gpu_render(sc, cam, deg, bg=(0.2, 0.1, 0.4), grad_weight=w)
        print("ok", m, P, W, H, float(color.mean()), flush=True)
    # equal depths: long tie runs
    sc, cam, rs = small_scene(P=1500, degree=0, W=96, H=64, eye=(0.0, 0.0, 3.0))
    flat = Scene(sc.means3D * torch.tensor([1.0, 1.0, 0.0]), sc.shs, sc.opacities, sc.scales, sc.rotations, 0)
    _cabi.set_option("bin_shift", -1)
    color, _, _ = gpu_render(flat, cam, 0, bg=(0.2, 0.1, 0.4), grad_weight=torch.rand(3, 64, 96))
    print("ok ties", m, float(color.mean()), flush=True)
    # a wall facing the camera: fat buckets -> the one-CTA radix path of the bucket sort
    if m["binning"] == 1:
        g = torch.Generator().manual_seed(3)
        P = 9000
        wall = Scene(torch.cat([torch.rand(P, 2, generator=g) * 2 - 1, torch.zeros(P, 1)], 1), sc.shs[:1].repeat(P, 1, 1),
                     torch.rand(P, 1, generator=g) * 0.3 + 0.05, torch.rand(P, 3, generator=g) * 0.03 + 0.01,
                     torch.tensor([[1.0, 0, 0, 0]]).repeat(P, 1), 0)
        _cabi.set_option("bin_shift", 3)
        color, _, _ = gpu_render(wall, cam, 0, bg=(0.2, 0.1, 0.4), grad_weight=torch.rand(3, 64, 96))
        print("ok wall", m, float(color.mean()), flush=True)
for k, v in dict(render=-1, gather=1, binning=-1, sort=1, sort_keys=32, bin_shift=-1, slab=0).items():
    _cabi.set_option(k, v)
from robosimgs_b200.losses import gs_loss
from robosimgs_b200.optim import FusedAdam
dev = torch.device("cuda:0")
x = torch.rand(3, 45, 70, device=dev, requires_grad=True)
gs_loss(x, torch.rand(3, 45, 70, device=dev)).backward()
opt = FusedAdam([x], lr=1e-2); opt.step()
torch.cuda.synchronize()
print("ok train neighbours", float(x.mean()))
# forward-only frames (B200GS_FORWARD_ONLY) with the 8-bit frame written by the compositing kernel (B200GS_OUT_RGB8):
# ragged widths (scalar epilogue) and multiples of 4 (packed words), both compositing kernel families
from robosimgs_b200 import GaussianRasterizer, rasterizer
from robosimgs_b200.scenes import settings_from_camera
for render in (0, 1):
    _cabi.set_option("render", render)
    for (W, H) in ((200, 120), (201, 119)):
        sc, cam, _ = small_scene(P=3000, degree=2, W=W, H=H)
        rs = settings_from_camera(cam, 2, bg=(0.2, 0.6, 0.4), device=dev)
        t = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        rasterizer._PAIR_HINTS.clear()
        with torch.no_grad():
            for attempt in range(2):
                out = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
                got, _, ticket = GaussianRasterizer(rs).forward_deferred(
                    t["means3D"], torch.zeros_like(t["means3D"]), t["opacities"], shs=t["shs"], scales=t["scales"],
                    rotations=t["rotations"], options=rasterizer.DeferOptions(rgb8=out))
                assert ticket.ok()
        print("ok rgb8", render, W, H, float(out.float().mean()), flush=True)
_cabi.set_option("render", -1)
