"""Small driver for ncu captures: a few forward+backward iterations of the C3 workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robosimgs_b200 import GaussianRasterizer
from robosimgs_b200.scenes import room_scene, room_target, settings_from_camera, mse_loss
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
sc, cam = room_scene()
rs = settings_from_camera(cam, 3, device=dev)
leaves = {k: getattr(sc, k).to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
target = room_target().to(dev)
r = GaussianRasterizer(rs)
for i in range(iters):
    color, radii = r(leaves["means3D"], m2d, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    mse_loss(color, target).backward()
torch.cuda.synchronize()
print("done", float(color.mean()))
