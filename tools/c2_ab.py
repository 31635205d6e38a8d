"""C2 (200k tabletop, 800x800, six reference views): per-stage times under the main options (dev tool)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robosimgs_b200 import GaussianRasterizer, _cabi
from robosimgs_b200.scenes import tabletop_scene, settings_from_camera
dev = torch.device("cuda:0")
sc, cams = tabletop_scene()
t = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2d = torch.zeros_like(t["means3D"])
for opts in (dict(render=-1), dict(render=1), dict(render=0)):
    _cabi.set_option("bin_shift", -1)
    for k, v in opts.items(): _cabi.set_option(k, v)
    for name, cam in cams.items() if isinstance(cams, dict) else enumerate(cams):
        rs = settings_from_camera(cam, 3, device=dev)
        r = GaussianRasterizer(rs)
        f = lambda: r(t["means3D"], m2d, t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
        with torch.no_grad():
            for _ in range(4): f()
            torch.cuda.synchronize(); _cabi.profile_enable(True); _cabi.profile_read(True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): f()
            b.record(); torch.cuda.synchronize()
            st = _cabi.profile_read(True); _cabi.profile_enable(False)
        print(opts, name, round(a.elapsed_time(b) / 10, 4), {k: round(v[0] / max(v[1], 1), 4) for k, v in st.items() if v[1]}, flush=True)
