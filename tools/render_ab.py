"""A/B of the compositing kernels on C3 (dev tool): one pixel per thread (render.cu) vs four pixels per
thread (render4.cu); opaque scene and the sparse worst case (opacity x0.03).  Writes gpurun_out/render_ab.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robosimgs_b200 import GaussianRasterizer, _cabi
from robosimgs_b200.scenes import room_scene, room_target, settings_from_camera, mse_loss

dev = torch.device("cuda:0")
sc, cam = room_scene()
rs = settings_from_camera(cam, 3, device=dev)
target = room_target().to(dev)
out = {}
modes = [int(m) for m in (sys.argv[1].split(",") if len(sys.argv) > 1 else ("0", "1"))]
for label, om, sm in (("opaque", 1.0, 1.0), ("sparse_x0.03", 0.03, 1.0), ("small_splats_x0.25", 1.0, 0.25)):
    res = {}
    for mode in modes:
        from robosimgs_b200 import rasterizer as _rz
        _rz._BIN_POLICY.clear(); _rz._PAIR_HINTS.clear()      # a fresh scene: let the bin-size policy look at it
        _cabi.set_option("render", mode)
        leaves = {k: getattr(sc, k).to(dev).clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        with torch.no_grad():
            leaves["opacities"].mul_(om)
            leaves["scales"].mul_(sm)
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
        r = GaussianRasterizer(rs)
        fwd = lambda: r(leaves["means3D"], m2d, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        def timed(grad, N):
            ctx = torch.enable_grad() if grad else torch.no_grad()
            with ctx:
                for _ in range(4):
                    c, _r = fwd()
                    if grad: mse_loss(c, target).backward()
                torch.cuda.synchronize(); _cabi.profile_enable(True); _cabi.profile_read(True)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(N):
                    c, _r = fwd()
                    if grad: mse_loss(c, target).backward()
                b.record(); torch.cuda.synchronize()
                st = _cabi.profile_read(True); _cabi.profile_enable(False)
            return a.elapsed_time(b) / N, {k: round(v[0] / max(v[1], 1), 4) for k, v in st.items() if v[1]}, c.detach()
        f_ms, f_st, img = timed(False, 30)
        t_ms, t_st, _ = timed(True, 15)
        for v in leaves.values(): v.grad = None
        c, _r = fwd(); mse_loss(c, target).backward(); torch.cuda.synchronize()
        res[mode] = dict(fwd_ms=round(f_ms, 4), fwd_stages=f_st, train_ms=round(t_ms, 4), train_stages=t_st)
        res[mode]["_img"] = img; res[mode]["_g"] = {k: v.grad.clone() for k, v in leaves.items()}
        print(label, "render mode", mode, json.dumps({k: v for k, v in res[mode].items() if not k.startswith("_")}), flush=True)
    if len(modes) == 2:
        a, b = res[modes[0]], res[modes[1]]
        mse = float(((a["_img"] - b["_img"]).double() ** 2).mean())
        cmp = {"psnr_between_modes": 10 * torch.log10(torch.tensor(1.0 / max(mse, 1e-30))).item()}
        for k in a["_g"]:
            cmp["grad_rel_" + k] = float((a["_g"][k] - b["_g"][k]).abs().max() / a["_g"][k].abs().max().clamp_min(1e-30))
        print(label, "cmp", json.dumps(cmp), flush=True)
        res["cmp"] = cmp
    out[label] = {str(k): ({kk: vv for kk, vv in v.items() if not kk.startswith("_")} if isinstance(v, dict) else v) for k, v in res.items()}
_cabi.set_option("render", -1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/render_ab.json", "w"), indent=1)
