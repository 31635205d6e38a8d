import sys, os, cProfile, pstats, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from robosimgs_b200 import GaussianRasterizer
from robosimgs_b200.scenes import cube_scene, settings_from_camera
dev = torch.device("cuda:0")
sc, cam = cube_scene()
rs = settings_from_camera(cam, 0, device=dev)
a = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2 = torch.zeros_like(a["means3D"])
def f():
    return GaussianRasterizer(rs)(a["means3D"], m2, a["opacities"], shs=a["shs"], scales=a["scales"], rotations=a["rotations"])
with torch.no_grad():
    for _ in range(20): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500): f()
    torch.cuda.synchronize()
    print("per call us", (time.perf_counter() - t0) / 500 * 1e6)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(500): f()
    pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
