"""Ad-hoc per-stage timing of the C3 workload (dev tool; not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robosimgs_b200 import GaussianRasterizer, _cabi, rasterizer
from robosimgs_b200.scenes import room_scene, room_target, settings_from_camera, mse_loss

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
sc, cam = room_scene(P)
rs = settings_from_camera(cam, 3, device=dev)
leaves = {k: getattr(sc, k).to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
target = room_target().to(dev)
r = GaussianRasterizer(rs)
def fwd():
    return r(leaves["means3D"], m2d, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
def run(label, N=20, grad=False, spec=True):
    rasterizer.SPECULATE_PAIR_CAPACITY = spec
    ctxm = torch.enable_grad() if grad else torch.no_grad()
    with ctxm:
        for i in range(3):
            color, radii = fwd()
            if grad: mse_loss(color, target).backward()
        torch.cuda.synchronize()
        _cabi.profile_enable(True); _cabi.profile_read(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cpu = 0.0
        ms0 = torch.cuda.memory_stats()
        per = []
        a.record()
        for i in range(N):
            t0 = time.perf_counter()
            color, radii = fwd()
            cpu += time.perf_counter() - t0
            per.append(round((time.perf_counter() - t0) * 1e3, 2))
            if grad: mse_loss(color, target).backward()
        b.record(); torch.cuda.synchronize()
        st = _cabi.profile_read(True); _cabi.profile_enable(False)
    ms = a.elapsed_time(b) / N
    stages = {k: round(v[0] / max(v[1], 1), 4) for k, v in st.items() if v[1]}
    ms1 = torch.cuda.memory_stats()
    print("   device_allocs", ms1["num_device_alloc"] - ms0["num_device_alloc"], "device_frees", ms1["num_device_free"] - ms0["num_device_free"], "reserved MB", ms1["reserved_bytes.all.current"] >> 20, "per-call ms", per[:12])
    print(f"{label}: {ms:.3f} ms/iter, cpu in forward() {cpu / N * 1e3:.3f} ms, gpu stage sum {sum(stages.values()):.3f} ms\n   {stages}", flush=True)
run("fwd sync-mode", spec=False)
run("fwd speculative", spec=True)
run("train speculative", grad=True)
run("fwd speculative again", spec=True)
run("train speculative again", grad=True)
