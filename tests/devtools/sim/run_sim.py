"""Work model of the compositing kernels on C3 (analysis only): python tests/devtools/sim/run_sim.py [P] [W] [H] [opacity_scale]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", ".."))
import torch
from oracle import gs_oracle
from robosimgs_b200.scenes import room_scene, settings_from_camera

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
osc = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
scene, cam = room_scene(P, 3, 3, W, H)
rs = settings_from_camera(cam, 3)
L = gs_oracle.lib()
dt = np.float32
prm = gs_oracle.make_params(rs, P, 16, dt)
means = scene.means3D.numpy(); shs = scene.shs.numpy(); opac = (scene.opacities.numpy().reshape(-1) * osc).astype(dt)
sc = scene.scales.numpy(); ro = scene.rotations.numpy()
radii = np.zeros(P, np.int32); xy = np.zeros((P, 2), dt); depths = np.zeros(P, dt)
cov3d = np.zeros((P, 6), dt); rgb = np.zeros((P, 3), dt); co = np.zeros((P, 4), dt); tt = np.zeros(P, np.int32); cl = np.zeros((P, 3), np.uint8)
p = lambda a: a.ctypes.data_as(C.c_void_p)
L.gso_preprocess_f32(C.byref(prm), p(means), p(shs), None, p(opac), p(sc), p(ro), None, p(radii), p(xy), p(depths), p(cov3d), p(rgb), p(co), p(tt), p(cl))
print("P_vis", int((radii > 0).sum()), "D_ref", int(tt.astype(np.int64).sum()))
sim = C.CDLL(os.path.join(os.path.dirname(__file__), "librender_sim.so"))
for bs in (3, 0):
    out = np.zeros(64, np.float64)
    t0 = time.time()
    sim.render_sim(P, p(xy), p(co), p(radii), p(depths), W, H, bs, 0, p(out))
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    print(f"bin_shift {bs}: pairs {out[0]:.0f} chunks {out[1]:.0f} ({out[1]/ntiles:.2f}/tile) walked {out[4]/ntiles:.0f}/tile "
          f"tile_surv {out[2]/ntiles:.0f}/tile px_contrib {out[3]:.3e} ({out[3]/ntiles/256:.1f}/px)  [{time.time()-t0:.1f}s]")
    names = ["8x4", "16x2", "8x8", "16x4", "16x8", "16x16"]
    for s, nm in enumerate(names):
        cb, sv, us = out[8 + 3 * s: 11 + 3 * s]
        area = int(nm.split("x")[0]) * int(nm.split("x")[1])
        print(f"   block {nm:6s} cull batches {cb:.3e} ({cb/ntiles:.1f}/tile)  block-survivor evals {sv:.3e} ({sv/ntiles:.1f}/tile)  useful px frac {us/max(sv*area,1):.3f}")
