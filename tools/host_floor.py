"""Host cost of one frame of the sweep (dev tool): SceneRenderer.submit/collect on a scene so small that the GPU is idle
most of the time -- what is left per frame is Python + graph launch + event handling.  Prints ms per frame for the
resident-frame path and the host-frame path, and a cProfile of the submit/collect loop."""
import cProfile, io, json, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robosimgs_b200.scenes import cube_scene
from robosimgs_b200.sweep import SceneRenderer
import bench

dev = torch.device("cuda:0")
sc, _ = cube_scene(P=2000, seed=3, degree=3)
tens = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
cams = bench.jittered_cameras(63)
blocks = torch.stack([torch.cat([c.viewmatrix.reshape(-1), c.projmatrix.reshape(-1), c.campos.reshape(-1)]) for c in cams]).to(dev)
bg = torch.zeros(3, device=dev)
out = {}
for name, host in (("resident", False), ("host_frames", True)):
    r = SceneRenderer(tens, 3, bg, 1080, 1920, streams=4 if not host else 6, graphs=True, host_frames=host)
    pend = []
    def sweep(n):
        for s in range(n):
            while len(pend) >= r.in_flight_limit():
                r.collect(pend.pop(0))
            pend.append(r.submit(cams[s % len(cams)], cam_block=None if host else blocks[s % len(cams)]))
        while pend:
            r.collect(pend.pop(0))
    with torch.no_grad():
        sweep(40)
        torch.cuda.synchronize()
        runs = []
        for _ in range(5):
            t0 = time.perf_counter(); sweep(200); torch.cuda.synchronize(); runs.append((time.perf_counter() - t0) / 200 * 1e3)
        out[name] = round(sorted(runs)[2], 4)
        if not host:
            pr = cProfile.Profile(); pr.enable(); sweep(400); pr.disable()
            s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
            print(s.getvalue()[:3500], file=sys.stderr)
    r.close()
print(json.dumps(out))
