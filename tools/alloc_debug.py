import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robosimgs_b200 import GaussianRasterizer, rasterizer
from robosimgs_b200.scenes import room_scene, settings_from_camera
dev = torch.device("cuda:0")
sc, cam = room_scene(1_000_000)
rs = settings_from_camera(cam, 3, device=dev)
t = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2d = torch.zeros_like(t["means3D"])
orig = rasterizer._POOL.acquire
def acq(key, n, device):
    before = sum(len(v) for v in rasterizer._POOL.free.values())
    r = orig(key, n, device)
    print(f"   acquire {key[2]} need {n>>20} MB got {r.numel()>>20} MB (pool had {before})")
    return r
rasterizer._POOL.acquire = acq
r = GaussianRasterizer(rs)
rasterizer.SPECULATE_PAIR_CAPACITY = False
with torch.no_grad():
    for i in range(6):
        a0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
        color, radii = r(t["means3D"], m2d, t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
        torch.cuda.synchronize()
        st = torch.cuda.memory_stats()
        print("iter", i, "device_allocs", st["num_device_alloc"] - a0, "reserved", st["reserved_bytes.all.current"] >> 20, "allocated", st["allocated_bytes.all.current"] >> 20)
