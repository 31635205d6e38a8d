"""ctypes front-end of the CPU oracle (oracle/gs_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs; never by the product package.

PARITY UNPINNED: the reference repo ships no rasterizer (anchor
/root/reference/README.md:75); this restates the published 3DGS algorithm.

The stage split mirrors the public pipeline:
``preprocess -> bin -> render`` and ``render_backward -> preprocess_backward``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgs_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/gs_oracle.c with the committed Makefile (gcc, OpenMP)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("gs_oracle.c", "gs_oracle_impl.h", "Makefile"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=sys.stderr)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.gso_bin_f32.restype = C.c_long
        _lib.gso_bin_f64.restype = C.c_long
    return _lib


def num_threads() -> int:
    return lib().gso_num_threads()


def set_num_threads(n: int) -> None:
    lib().gso_set_num_threads(int(n))


def _params_struct(real):
    class P(C.Structure):
        _fields_ = [
            ("P", C.c_int), ("sh_degree", C.c_int), ("M", C.c_int), ("H", C.c_int), ("W", C.c_int),
            ("tanfovx", real), ("tanfovy", real), ("scale_modifier", real), ("denom_eps", real),
            ("near_plane", real),
            ("bg", real * 3), ("view", real * 16), ("proj", real * 16), ("campos", real * 3),
        ]
    return P


_PF32 = _params_struct(C.c_float)
_PF64 = _params_struct(C.c_double)


def _np(x, dt):
    """torch tensor / array-like / None -> contiguous numpy array of dtype dt (or None)."""
    if x is None:
        return None
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dt)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_params(settings, P, M, dtype=np.float32, denom_eps=1e-7, near_plane=0.2):
    """settings: any object with the GaussianRasterizationSettings field names."""
    S = _PF32 if dtype == np.float32 else _PF64
    p = S()
    p.P, p.sh_degree, p.M = int(P), int(settings.sh_degree), int(M)
    p.H, p.W = int(settings.image_height), int(settings.image_width)
    p.tanfovx, p.tanfovy = float(settings.tanfovx), float(settings.tanfovy)
    p.scale_modifier = float(settings.scale_modifier)
    p.denom_eps = float(denom_eps)
    p.near_plane = float(near_plane)
    bg = _np(settings.bg, np.float64).reshape(-1)
    vm = _np(settings.viewmatrix, np.float64).reshape(-1)
    pm = _np(settings.projmatrix, np.float64).reshape(-1)
    cp = _np(settings.campos, np.float64).reshape(-1)
    for i in range(3):
        p.bg[i] = bg[i]
        p.campos[i] = cp[i]
    for i in range(16):
        p.view[i] = vm[i]
        p.proj[i] = pm[i]
    return p


def forward(settings, means3D, opacities, shs=None, colors_precomp=None, scales=None,
            rotations=None, cov3D_precomp=None, dtype=np.float32, denom_eps=1e-7, near_plane=0.2):
    """Full oracle forward.  Returns a namespace with the image, radii and every intermediate."""
    L = lib()
    sfx = "_f32" if dtype == np.float32 else "_f64"
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    means = _np(means3D, dtype).reshape(-1, 3)
    P = means.shape[0]
    shs_ = _np(shs, dtype)
    M = 0 if shs_ is None else (shs_.shape[1] if shs_.ndim == 3 else shs_.reshape(P, -1, 3).shape[1])
    prm = make_params(settings, P, M, dtype, denom_eps, near_plane)
    H, W = prm.H, prm.W
    st = SimpleNamespace(dtype=dtype, P=P, M=M, H=H, W=W, params=prm, sfx=sfx)
    st.means, st.shs = means, shs_
    st.colors_precomp = _np(colors_precomp, dtype)
    st.opac = _np(opacities, dtype).reshape(-1)
    st.scales, st.rots = _np(scales, dtype), _np(rotations, dtype)
    st.cov3d_precomp = _np(cov3D_precomp, dtype)
    st.radii = np.zeros(P, np.int32)
    st.xy = np.zeros((P, 2), dtype)
    st.depths = np.zeros(P, dtype)
    st.cov3d = np.zeros((P, 6), dtype)
    st.rgb = np.zeros((P, 3), dtype)
    st.conic_opacity = np.zeros((P, 4), dtype)
    st.tiles_touched = np.zeros(P, np.int32)
    st.clamped = np.zeros((P, 3), np.uint8)
    getattr(L, "gso_preprocess" + sfx)(
        C.byref(prm), _ptr(means), _ptr(shs_), _ptr(st.colors_precomp), _ptr(st.opac),
        _ptr(st.scales), _ptr(st.rots), _ptr(st.cov3d_precomp), _ptr(st.radii), _ptr(st.xy),
        _ptr(st.depths), _ptr(st.cov3d), _ptr(st.rgb), _ptr(st.conic_opacity),
        _ptr(st.tiles_touched), _ptr(st.clamped))
    D = int(st.tiles_touched.astype(np.int64).sum())
    st.num_rendered = D
    gx, gy = (W + 15) // 16, (H + 15) // 16
    st.point_list = np.zeros(max(D, 1), np.int32)
    st.ranges = np.zeros((gx * gy, 2), np.int32)
    n = getattr(L, "gso_bin" + sfx)(C.byref(prm), _ptr(st.radii), _ptr(st.xy), _ptr(st.depths),
                                    C.c_long(D), _ptr(st.point_list), _ptr(st.ranges))
    assert n == D, (n, D)
    st.color = np.zeros((3, H, W), dtype)
    st.final_T = np.zeros((H, W), dtype)
    st.n_contrib = np.zeros((H, W), np.int32)
    getattr(L, "gso_render" + sfx)(C.byref(prm), _ptr(st.ranges), _ptr(st.point_list), _ptr(st.xy),
                                   _ptr(st.rgb), _ptr(st.conic_opacity), _ptr(st.color),
                                   _ptr(st.final_T), _ptr(st.n_contrib))
    return st


def preprocess_only(settings, means3D, opacities, shs, scales, rotations, dtype=np.float32):
    """Steps 1-9 only (the 'Python/CPU preprocess path' north_star asks to be timed)."""
    L = lib()
    sfx = "_f32" if dtype == np.float32 else "_f64"
    means = _np(means3D, dtype).reshape(-1, 3)
    P = means.shape[0]
    shs_ = _np(shs, dtype)
    M = shs_.reshape(P, -1, 3).shape[1]
    prm = make_params(settings, P, M, dtype)
    opac, sc, ro = _np(opacities, dtype).reshape(-1), _np(scales, dtype), _np(rotations, dtype)
    radii = np.zeros(P, np.int32); xy = np.zeros((P, 2), dtype); depths = np.zeros(P, dtype)
    cov3d = np.zeros((P, 6), dtype); rgb = np.zeros((P, 3), dtype)
    co = np.zeros((P, 4), dtype); tt = np.zeros(P, np.int32); cl = np.zeros((P, 3), np.uint8)

    def run():
        getattr(L, "gso_preprocess" + sfx)(
            C.byref(prm), _ptr(means), _ptr(shs_), None, _ptr(opac), _ptr(sc), _ptr(ro), None,
            _ptr(radii), _ptr(xy), _ptr(depths), _ptr(cov3d), _ptr(rgb), _ptr(co), _ptr(tt), _ptr(cl))
        return radii, tt
    return run


def backward(st, dL_dcolor):
    """Oracle backward for a state returned by forward().  Returns a namespace of gradients."""
    L = lib()
    dt, P, sfx, prm = st.dtype, st.P, st.sfx, st.params
    dpix = _np(dL_dcolor, dt).reshape(3, st.H, st.W)
    g = SimpleNamespace()
    g.mean2D = np.zeros((P, 2), dt)
    g.conic = np.zeros((P, 3), dt)
    g.opacity = np.zeros(P, dt)
    g.color = np.zeros((P, 3), dt)
    getattr(L, "gso_render_backward" + sfx)(
        C.byref(prm), _ptr(st.ranges), _ptr(st.point_list), _ptr(st.xy), _ptr(st.conic_opacity),
        _ptr(st.rgb), _ptr(st.final_T), _ptr(st.n_contrib), _ptr(dpix), _ptr(g.mean2D),
        _ptr(g.conic), _ptr(g.opacity), _ptr(g.color))
    g.means3D = np.zeros((P, 3), dt)
    g.shs = None if st.shs is None else np.zeros((P, st.M, 3), dt)
    g.scales = None if st.scales is None else np.zeros((P, 3), dt)
    g.rotations = None if st.rots is None else np.zeros((P, 4), dt)
    g.cov3D = np.zeros((P, 6), dt)
    getattr(L, "gso_preprocess_backward" + sfx)(
        C.byref(prm), _ptr(st.means), _ptr(st.shs), _ptr(st.scales), _ptr(st.rots),
        _ptr(st.cov3d), _ptr(st.radii), _ptr(st.clamped), _ptr(g.mean2D), _ptr(g.conic),
        _ptr(g.color), _ptr(g.means3D), _ptr(g.shs), _ptr(g.scales), _ptr(g.rotations),
        _ptr(g.cov3D))
    g.colors_precomp = g.color if st.colors_precomp is not None else None
    g.opacities = g.opacity.reshape(P, 1)
    g.means2D = np.concatenate([g.mean2D, np.zeros((P, 1), dt)], axis=1)
    return g


def mark_visible(settings, positions, dtype=np.float32):
    means = _np(positions, dtype).reshape(-1, 3)
    prm = make_params(settings, means.shape[0], 0, dtype)
    out = np.zeros(means.shape[0], np.uint8)
    sfx = "_f32" if dtype == np.float32 else "_f64"
    getattr(lib(), "gso_mark_visible" + sfx)(C.byref(prm), _ptr(means), _ptr(out))
    return out.astype(bool)
