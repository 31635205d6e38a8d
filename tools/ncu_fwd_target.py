"""Small driver for ncu captures: a few forward frames of the C3 workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robosimgs_b200 import GaussianRasterizer
from robosimgs_b200.scenes import room_scene, settings_from_camera
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
sc, cam = room_scene()
rs = settings_from_camera(cam, 3, device=dev)
t = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2d = torch.zeros_like(t["means3D"])
r = GaussianRasterizer(rs)
with torch.no_grad():
    for i in range(iters):
        color, radii = r(t["means3D"], m2d, t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
torch.cuda.synchronize()
print("done", float(color.mean()))
