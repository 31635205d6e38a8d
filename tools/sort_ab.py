import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from robosimgs_b200 import GaussianRasterizer, _cabi
from robosimgs_b200.scenes import cube_scene, tabletop_scene, room_scene, settings_from_camera
dev = torch.device("cuda:0")
def t(sc, cam, deg, label):
    rs = settings_from_camera(cam, deg, device=dev)
    a = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    m2 = torch.zeros_like(a["means3D"])
    f = lambda: GaussianRasterizer(rs)(a["means3D"], m2, a["opacities"], shs=a["shs"], scales=a["scales"], rotations=a["rotations"])
    for mode in (0, 1):
        _cabi.set_option("sort", mode)
        with torch.no_grad():
            for _ in range(5): f()
            torch.cuda.synchronize(); _cabi.profile_enable(True); _cabi.profile_read(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): f()
            e1.record(); torch.cuda.synchronize()
            st = _cabi.profile_read(True); _cabi.profile_enable(False)
        print(label, "sort", "coop" if mode else "cub", "frame ms", round(e0.elapsed_time(e1) / 20, 4), "pair_sort ms", round(st["pair_sort"][0] / max(st["pair_sort"][1], 1), 4), flush=True)
sc, cam = cube_scene(); t(sc, cam, 0, "C1")
sc, cams = tabletop_scene(); t(sc, cams["top"], 3, "C2")
sc, cam = room_scene(); t(sc, cam, 3, "C3")
