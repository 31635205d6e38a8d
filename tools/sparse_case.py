"""Worst case for coarse bins: nothing saturates, every CTA walks its whole bin list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robosimgs_b200 import GaussianRasterizer, _cabi
from robosimgs_b200.scenes import room_scene, settings_from_camera
dev = torch.device("cuda:0")
sc, cam = room_scene()
rs = settings_from_camera(cam, 3, device=dev)
for label, om in (("normal", 1.0), ("opacity x0.03", 0.03)):
    t = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    t["opacities"] = t["opacities"] * om
    m2d = torch.zeros_like(t["means3D"])
    m3 = t["means3D"].clone().requires_grad_(True)
    r = GaussianRasterizer(rs)
    c, _ = r(m3, m2d, t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
    D = c.grad_fn.num_rendered
    with torch.no_grad():
        for i in range(5): r(t["means3D"], m2d, t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
        torch.cuda.synchronize(); _cabi.profile_enable(True); _cabi.profile_read(True)
        for i in range(10): r(t["means3D"], m2d, t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
        torch.cuda.synchronize(); st = _cabi.profile_read(True); _cabi.profile_enable(False)
    print(label, "D", D, {k: round(v[0] / max(v[1], 1), 4) for k, v in st.items() if v[1]}, flush=True)
