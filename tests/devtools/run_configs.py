"""Runs every BASELINE.json config on one GPU and writes gpurun_out/configs_r2.json
(GPU ms, Mpix/s, pair counts; PSNR / gradient error against the oracle where the CPU oracle is
affordable).  Dev/documentation tool -- bench.py is the contract benchmark.  Lives under tests/ because it uses
the oracle as its checker (only tests/, smoke() and bench.py's CPU arms may touch oracle/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from robosimgs_b200 import GaussianRasterizer, _cabi, compositor as cp, export_rgb8
from robosimgs_b200.scenes import (cube_scene, tabletop_scene, room_scene, room_target, sweep_scene,
                                   settings_from_camera, mse_loss)
from oracle import gs_oracle
from helpers import psnr, max_rel_err

dev = torch.device("cuda:0")
out = {}

def gpu_time(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def on_dev(sc, grad=False):
    return {k: getattr(sc, k).to(dev).requires_grad_(grad) for k in ("means3D", "shs", "opacities", "scales", "rotations")}

def fwd(t, rs, m2d):
    return GaussianRasterizer(rs)(t["means3D"], m2d, t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"])

# ---- C1 ----
sc, cam = cube_scene()
rs, rs_cpu = settings_from_camera(cam, 0, device=dev), settings_from_camera(cam, 0)
t = on_dev(sc, True); m2d = torch.zeros_like(t["means3D"], requires_grad=True)
w = torch.rand(3, 256, 256, generator=torch.Generator().manual_seed(1))
color, radii = fwd(t, rs, m2d); (color * w.to(dev)).sum().backward()
st = gs_oracle.forward(rs_cpu, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations, dtype=np.float64)
ref = gs_oracle.backward(st, w.numpy())
gerr = max(max_rel_err(t[k].grad.cpu().numpy().reshape(getattr(ref, k).shape), getattr(ref, k)) for k in ("means3D", "shs", "opacities", "scales", "rotations"))
with torch.no_grad():
    ms = gpu_time(lambda: fwd(t, rs, m2d))
def c1_train():
    c, _ = fwd(t, rs, m2d); (c * c).mean().backward()
ms_tr = gpu_time(c1_train)
t0 = time.perf_counter(); s32 = gs_oracle.forward(rs_cpu, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations); gs_oracle.backward(s32, w.numpy()); cpu_s = time.perf_counter() - t0
out["C1"] = dict(P=sc.P, image=[256, 256], fwd_ms=ms, mpix_s=256 * 256 / ms / 1e3, train_ms=ms_tr, D=int(color.grad_fn.num_rendered), D_ref=int(st.num_rendered),
                 psnr=psnr(color.detach().cpu().numpy(), st.color), grad_max_rel_err=gerr, cpu_fwd_bwd_s=cpu_s)
print("C1", out["C1"], flush=True)

# ---- C2 ----
sc, cams = tabletop_scene()
t = on_dev(sc); m2d = torch.zeros_like(t["means3D"])
res = {}
with torch.no_grad():
    for name, cam in cams.items():
        rs = settings_from_camera(cam, 3, device=dev)
        ms = gpu_time(lambda: fwd(t, rs, m2d), n=10, warm=3)
        res[name] = ms
    cam = cams["top"]; rs = settings_from_camera(cam, 3, device=dev)
    color, _ = fwd(t, rs, m2d)
t0 = time.perf_counter(); st = gs_oracle.forward(settings_from_camera(cam, 3), sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations); cpu_s = time.perf_counter() - t0
out["C2"] = dict(P=sc.P, image=[800, 800], fwd_ms_per_view=res, mpix_s=0.64 / (sum(res.values()) / 6) * 1e3, psnr_top=psnr(color.cpu().numpy(), st.color), cpu_fwd_s=cpu_s, D_ref=int(st.num_rendered))
print("C2", out["C2"], flush=True)

# ---- C3 ----
sc, cam = room_scene()
rs = settings_from_camera(cam, 3, device=dev)
t = on_dev(sc, True); m2d = torch.zeros_like(t["means3D"], requires_grad=True)
target = room_target().to(dev)
with torch.no_grad():
    ms = gpu_time(lambda: fwd(t, rs, m2d))
def c3_train():
    c, _ = fwd(t, rs, m2d); mse_loss(c, target).backward()
ms_tr = gpu_time(c3_train)
c, r = fwd(t, rs, m2d)
out["C3"] = dict(P=sc.P, image=[1920, 1080], fwd_ms=ms, mpix_s=1920 * 1080 / ms / 1e3, train_ms=ms_tr, iters_s=1e3 / ms_tr, D=int(c.grad_fn.num_rendered), P_vis=int((r > 0).sum()))
print("C3", out["C3"], flush=True)
del t, m2d, c, r

# ---- C4 (single GPU share: 8 of the 64 cameras), through the datagen renderer (host frames) ----
from robosimgs_b200.sweep import SceneRenderer
sc, cams = sweep_scene()
t = on_dev(sc)
mine = cams[0::8]
r = SceneRenderer(t, 3, torch.zeros(3, device=dev), 1080, 1920, streams=4)
def c4_pass(n):
    pend = []
    for k in range(n):
        for cam in mine:
            while len(pend) >= r.in_flight_limit():
                r.collect(pend.pop(0))
            pend.append(r.submit(cam))
    while pend:
        r.collect(pend.pop(0))
with torch.no_grad():
    c4_pass(3)
    torch.cuda.synchronize(); r.redone = 0; t0 = time.perf_counter()
    c4_pass(8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["C4_1gpu_share"] = dict(P=sc.P, cameras=len(mine), passes=8, frames_s=8 * len(mine) / dt, ms_per_frame=dt / (8 * len(mine)) * 1e3,
                            frames_rendered_twice=r.redone, what="SceneRenderer, 4 frames in flight, 8-bit frames to pinned host memory")
print("C4", out["C4_1gpu_share"], flush=True)
del t, r

# ---- C5: room + the reference's openbox object (fixture samples), lid about the fixture hinge axis ----
bg, cam = room_scene()
ob = np.load(os.path.join(ROOT, "tests", "golden", "openbox_surface_samples.npz"))
obj, link_ids = cp.object_from_surface_samples(ob["body_pts"], ob["body_nrm"].astype(np.float32), ob["lid_pts"], ob["lid_nrm"].astype(np.float32))
art = cp.ArticulatedScene(bg, obj, link_ids, dev)
rs = settings_from_camera(cam, 3, device=dev)
axis, scale = ob["axis"], 0.1
base_q = cp.axis_angle_quat((1, 0, 0), -np.pi / 2); base_t = (1.2, 0.6, -1.3)
def frame(f):
    T0, q0 = cp.revolute_link_pose(axis, (0, 0, 0), 0.0, base_q=base_q, base_t=base_t, scale=scale)
    T1, q1 = cp.revolute_link_pose(axis, (0, 0, 0), cp.lid_angle(f), base_q=base_q, base_t=base_t, scale=scale)
    art.set_link_poses(np.stack([T0, T1]), np.stack([q0, q1]), scale=scale)
    return export_rgb8(art.render(rs)[0])
for f in range(5): frame(f)
torch.cuda.synchronize(); t0 = time.perf_counter()
for f in range(120): img = frame(f)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["C5"] = dict(P=art.P_bg + art.P_obj, frames=120, frames_s=120 / dt, ms_per_frame=dt / 120 * 1e3, obj_visible=int((art.render(rs)[1][art.P_bg:] > 0).sum()),
                 object="openbox_output/urdf body_centered.glb + lid_centered.glb surface samples (tests/golden), hinge axis of metadata.json")
print("C5", out["C5"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_r2.json"), "w"), indent=1)
