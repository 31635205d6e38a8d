"""Small fwd+bwd cases for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import small_scene, gpu_render
from robosimgs_b200 import _cabi
for gather in (1, 0):
    _cabi.set_option("gather", gather)
    for (P, W, H, deg, shift) in ((3000, 200, 136, 3, -1), (6000, 64, 48, 0, 0), (500, 37, 23, 1, 2)):
        _cabi.set_option("bin_shift", shift)
        sc, cam, rs = small_scene(P=P, degree=deg, W=W, H=H)
        if P == 6000:
            sc.opacities.mul_(0.15)
        w = torch.rand(3, H, W)
        color, radii, grads = gpu_render(sc, cam, deg, bg=(0.2, 0.1, 0.4), grad_weight=w)
        print("ok", gather, P, W, H, float(color.mean()), flush=True)
