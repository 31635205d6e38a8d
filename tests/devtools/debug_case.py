import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import small_scene, gpu_render, psnr
from oracle import gs_oracle
for P, opm, fov in [(6000, 0.15, 35.0), (6000, 0.15, 60.0), (1500, 0.15, 35.0), (6000, 1.0, 35.0), (3000, 0.15, 35.0)]:
    sc, cam, rs = small_scene(P=P, degree=0, W=64, H=48, big=0, fov=fov)
    sc.opacities.mul_(opm)
    color, radii, _ = gpu_render(sc, cam, 0)
    st = gs_oracle.forward(rs, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations, dtype=np.float64)
    L = st.ranges[:,1]-st.ranges[:,0]
    print(f"P={P} opm={opm} fov={fov}: psnr {psnr(color, st.color):.1f} radii mismatch {(radii!=st.radii).sum()} maxlist {L.max()} D_ref {st.num_rendered}")
    d = np.abs(color - st.color).max(0)
    for ty in range(3):
        print("   tile maxdiff:", " ".join(f"{d[ty*16:(ty+1)*16, tx*16:(tx+1)*16].max():.4f}" for tx in range(4)), "  lists:", L.reshape(3,4)[ty])
