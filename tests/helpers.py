"""Shared helpers for the parity tests."""
import numpy as np
import torch

from robosimgs_b200.cameras import camera_look_at
from robosimgs_b200.scenes import Scene, cube_scene, settings_from_camera


def psnr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mse = float(((a - b) ** 2).mean())
    return 10.0 * np.log10(1.0 / max(mse, 1e-30))


def max_rel_err(got, ref):
    """Gradient tolerance metric used throughout: max |got - ref| / max |ref| per tensor
    (north_star: grad max-rel-err < 1e-3)."""
    got = np.asarray(got, np.float64).reshape(-1)
    ref = np.asarray(ref, np.float64).reshape(-1)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def component_max_rel_err(got, ref):
    """Second gradient metric, per COMPONENT: for a tensor [P, ...] the per-tensor metric above is evaluated
    separately for every trailing index (each SH coefficient and colour channel, each axis of a scale, ...),
    so an error in a small-magnitude component (an SH degree-3 coefficient next to DC) cannot hide behind
    the tensor's largest element.  Returns the worst component."""
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64).reshape(ref.shape)
    P = ref.shape[0]
    d = np.abs(got - ref).reshape(P, -1).max(axis=0)
    s = np.abs(ref).reshape(P, -1).max(axis=0)
    ok = s > 0
    return float((d[ok] / s[ok]).max()) if ok.any() else 0.0


def elementwise_violations(got, ref, rtol=1e-3, atol_frac=1e-4):
    """Element-wise gradient bound: fraction of elements with |got - ref| > rtol |ref| + atol_frac * (max |ref| of
    the element's component).  fp32 accumulation over thousands of pixel terms leaves an absolute error that
    scales with the component, not with the element, hence the small absolute floor."""
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64).reshape(ref.shape)
    P = ref.shape[0]
    r2, g2 = ref.reshape(P, -1), got.reshape(P, -1)
    floor = atol_frac * np.abs(r2).max(axis=0, keepdims=True)
    bad = np.abs(g2 - r2) > rtol * np.abs(r2) + floor
    return float(bad.mean())


def small_scene(P=400, seed=5, degree=3, W=72, H=56, big=20, fov=60.0, eye=(0.5, 0.3, 2.5),
                bg=(0.2, 0.1, 0.4), scale_modifier=1.0, opacity_boost=0.0):
    sc, _ = cube_scene(P=P, seed=seed, degree=degree)
    if big:
        sc.scales[:big] *= 8
    if opacity_boost:
        sc.opacities.copy_(torch.sigmoid(torch.logit(sc.opacities) + opacity_boost))
    cam = camera_look_at(eye, (0, 0, 0), (0, 1, 0), fov, W, H)
    rs = settings_from_camera(cam, degree, bg=bg, scale_modifier=scale_modifier)
    return sc, cam, rs


def gpu_render(scene: Scene, cam, degree, bg=(0, 0, 0), scale_modifier=1.0, grad_weight=None,
               colors_precomp=None, cov3D_precomp=None, debug=False):
    """Run the CUDA path through the public operator; returns (color, radii, grads dict)."""
    from robosimgs_b200 import GaussianRasterizer
    dev = torch.device("cuda:0")
    rs = settings_from_camera(cam, degree, bg=bg, scale_modifier=scale_modifier, device=dev)
    if debug:
        rs = rs._replace(debug=True)
    need_grad = grad_weight is not None
    leaf = lambda t: t.to(dev).clone().requires_grad_(need_grad)
    means3D, opac = leaf(scene.means3D), leaf(scene.opacities)
    means2D = torch.zeros_like(means3D, requires_grad=need_grad)
    kw, leaves = {}, {"means3D": means3D, "opacities": opac, "means2D": means2D}
    if colors_precomp is not None:
        kw["colors_precomp"] = leaves["colors_precomp"] = leaf(colors_precomp)
    else:
        kw["shs"] = leaves["shs"] = leaf(scene.shs)
    if cov3D_precomp is not None:
        kw["cov3D_precomp"] = leaves["cov3D_precomp"] = leaf(cov3D_precomp)
    else:
        kw["scales"] = leaves["scales"] = leaf(scene.scales)
        kw["rotations"] = leaves["rotations"] = leaf(scene.rotations)
    color, radii = GaussianRasterizer(rs)(means3D, means2D, opac, **kw)
    grads = {}
    if need_grad:
        (color * grad_weight.to(dev)).sum().backward()
        grads = {k: v.grad.detach().cpu().numpy() for k, v in leaves.items()}
    torch.cuda.synchronize()
    return color.detach().cpu().numpy(), radii.cpu().numpy(), grads
