"""Diagnose the single-stream pass (dev tool): per run of 60 frames the device time, host time, library launches,
new CUDA allocations and the pair counts, to explain outliers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from robosimgs_b200 import GaussianRasterizer, _cabi, rasterizer
from robosimgs_b200.scenes import room_scene, settings_from_camera
dev = torch.device("cuda:0")
sc, _ = room_scene()
tens = {k: getattr(sc, k).to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2d = torch.zeros_like(tens["means3D"])
cams = bench.jittered_cameras(63)
rs = [settings_from_camera(c, 3, device=dev) for c in cams]
def frame(s):
    return GaussianRasterizer(rs[s])(tens["means3D"], m2d, tens["opacities"], shs=tens["shs"], scales=tens["scales"], rotations=tens["rotations"])
with torch.no_grad():
    for s in range(8): frame(s)
    for run in range(8):
        torch.cuda.synchronize()
        st0 = torch.cuda.memory_stats()
        _cabi.launch_count(reset=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        host = []
        for s in range(3, 63):
            h0 = time.perf_counter(); frame(s); host.append(time.perf_counter() - h0)
        b.record(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
        st1 = torch.cuda.memory_stats()
        key = (0, 1_000_000, 1080, 1920)
        print(f"run {run}: device {a.elapsed_time(b)/60:.4f} ms/frame, wall {wall/60*1e3:.4f}, host max {max(host)*1e3:.3f} med {sorted(host)[30]*1e3:.3f} ms, "
              f"launches {_cabi.launch_count()/60:.1f}/frame, cudaMallocs {st1['num_device_alloc']-st0['num_device_alloc']}, "
              f"hint {rasterizer._PAIR_HINTS.get(key)}, shift {rasterizer._BIN_POLICY.get(key)}", flush=True)
