"""A/B of the binning pipelines on C3 (dev tool): bucketed vs global sort, several bin sizes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robosimgs_b200 import GaussianRasterizer, _cabi
from robosimgs_b200.scenes import room_scene, room_target, settings_from_camera
from robosimgs_b200.losses import mse_loss

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
sc, cam = room_scene(P)
rs = settings_from_camera(cam, 3, device=dev)
target = room_target().to(dev)
leaves = {k: getattr(sc, k).to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
r = GaussianRasterizer(rs)
fwd = lambda: r(leaves["means3D"], m2d, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
ref = None
for binning, shift in ((0, -1), (1, -1), (1, 3), (1, 1), (1, 0)):
    _cabi.set_option("binning", binning); _cabi.set_option("bin_shift", shift)
    def timed(grad, N):
        ctx = torch.enable_grad() if grad else torch.no_grad()
        with ctx:
            for _ in range(4):
                c, _r = fwd()
                if grad: mse_loss(c, target).backward()
            torch.cuda.synchronize(); _cabi.profile_enable(True); _cabi.profile_read(True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(N):
                c, _r = fwd()
                if grad: mse_loss(c, target).backward()
            b.record(); torch.cuda.synchronize()
            st = _cabi.profile_read(True); _cabi.profile_enable(False)
        return a.elapsed_time(b) / N, {k: round(v[0] / max(v[1], 1), 4) for k, v in st.items() if v[1]}, c
    f_ms, f_st, img = timed(False, 30)
    t_ms, t_st, c = timed(True, 15)
    D = c.grad_fn.num_rendered
    if ref is None: ref = img.clone()
    print(json.dumps(dict(binning=binning, shift=shift, D=D, fwd_ms=round(f_ms, 4), train_ms=round(t_ms, 4), fwd=f_st,
                          train=t_st, identical=bool(torch.equal(ref, img)))), flush=True)
_cabi.set_option("binning", -1); _cabi.set_option("bin_shift", -1)
