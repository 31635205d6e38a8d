"""Host-side cost of the end-to-end frame loop (dev tool): cProfile of the bench's e2e step."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robosimgs_b200 import GaussianRasterizer, export_rgb8
from robosimgs_b200.rasterizer import GaussianRasterizationSettings
from robosimgs_b200.scenes import room_scene
from robosimgs_b200.sweep import FrameStreams
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import jittered_cameras
dev = torch.device("cuda:0")
sc, _ = room_scene()
tens = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
means2D = torch.zeros_like(tens["means3D"])
cams = jittered_cameras(64)
packs = [torch.cat([c.viewmatrix.reshape(-1), c.projmatrix.reshape(-1), c.campos.reshape(-1)]).pin_memory() for c in cams]
bg = torch.zeros(3, device=dev)
NS, NBUF = 2, 4
host = [torch.empty((1080, 1920, 3), dtype=torch.uint8).pin_memory() for _ in range(NBUF)]
done = [torch.cuda.Event() for _ in range(NBUF)]
tick = [None] * NBUF
fs = FrameStreams(dev, NS)
def step(s):
    with fs.next():
        c = cams[s % 64]; d = packs[s % 64].to(dev, non_blocking=True)
        rs = GaussianRasterizationSettings(c.image_height, c.image_width, c.tanfovx, c.tanfovy, bg, 1.0,
                                           d[0:16].view(4, 4), d[16:32].view(4, 4), 3, d[32:35], False, False)
        col, _, t = GaussianRasterizer(rs).forward_deferred(tens["means3D"], means2D, tens["opacities"], shs=tens["shs"],
                                                            scales=tens["scales"], rotations=tens["rotations"])
        col = export_rgb8(col)
        host[s % NBUF].copy_(col, non_blocking=True)
        done[s % NBUF].record(); tick[s % NBUF] = t
    if s >= NS:
        done[(s - NS) % NBUF].synchronize(); assert tick[(s - NS) % NBUF].ok()
with torch.no_grad():
    for s in range(20): step(s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(200): step(s)
    torch.cuda.synchronize()
    print("ms per frame", (time.perf_counter() - t0) / 200 * 1e3)
    pr = cProfile.Profile(); pr.enable()
    for s in range(300): step(s)
    pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
