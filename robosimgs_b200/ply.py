"""3DGS .ply reader / writer.

The reference names "a .ply file" as the hand-off between background reconstruction and everything
downstream (/root/reference/README.md:75).  Format (the public 3DGS layout): binary little-endian,
one `vertex` element of float32 properties
    x y z nx ny nz f_dc_0..2 f_rest_0..(3*n_rest-1) opacity scale_0..2 rot_0..3
stored PRE-activation (opacity = logit, scale = log, rot un-normalised wxyz, f_rest channel-major).
Reading = numpy parse of the header + one H2D copy of the vertex block + ONE fused CUDA kernel
(libb200gs: b200gs_ply_activate) that applies sigmoid / exp / normalise and repacks SH.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from .scenes import Scene


def _property_names(n_rest: int):
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    names += [f"f_rest_{i}" for i in range(3 * n_rest)]
    names += ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    return names


def write_gaussian_ply(path: str, means3D, shs, opacities, scales, rotations) -> None:
    """Write post-activation tensors (the GaussianRasterizer.forward layout) as a 3DGS .ply:
    inverse activations are applied (logit / log); quaternions are written as given."""
    means = np.asarray(torch.as_tensor(means3D).detach().cpu(), np.float32)
    sh = np.asarray(torch.as_tensor(shs).detach().cpu(), np.float32)
    P, M = sh.shape[0], sh.shape[1]
    n_rest = M - 1
    op = np.asarray(torch.as_tensor(opacities).detach().cpu(), np.float64).reshape(P)
    op = np.clip(op, 1e-7, 1 - 1e-7)
    sc = np.asarray(torch.as_tensor(scales).detach().cpu(), np.float64)
    rot = np.asarray(torch.as_tensor(rotations).detach().cpu(), np.float32)
    cols = [means, np.zeros((P, 3), np.float32), sh[:, 0, :],
            sh[:, 1:, :].transpose(0, 2, 1).reshape(P, 3 * n_rest),          # channel-major
            np.log(op / (1 - op)).astype(np.float32)[:, None], np.log(sc).astype(np.float32), rot]
    data = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype="<f4")
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {P}\n"
    header += "".join(f"property float {n}\n" for n in _property_names(n_rest)) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(data.tobytes())


def read_gaussian_ply(path: str):
    """Parse a binary little-endian 3DGS .ply.  Returns (vertices float32 [P, stride], layout dict):
    the raw pre-activation records and the float offsets of each property group."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, P, props, in_vertex = None, None, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    P = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] not in ("float", "float32"):
                    raise ValueError(f"{path}: non-float vertex property {tok[2]} ({tok[1]})")
                props.append(tok[2])
            elif tok[0] == "end_header":
                break
        if fmt != "binary_little_endian" or P is None:
            raise ValueError(f"{path}: only binary_little_endian PLY with a vertex element is supported")
        stride = len(props)
        raw = np.frombuffer(f.read(P * stride * 4), dtype="<f4").copy()
        if raw.size != P * stride:
            raise ValueError(f"{path}: truncated vertex data")
    idx = {n: i for i, n in enumerate(props)}
    n_rest3 = sum(1 for n in props if n.startswith("f_rest_"))
    if n_rest3 % 3:
        raise ValueError(f"{path}: f_rest count {n_rest3} is not a multiple of 3")
    for group in (("x", "y", "z"), ("f_dc_0", "f_dc_1", "f_dc_2"), ("scale_0", "scale_1", "scale_2"),
                  ("rot_0", "rot_1", "rot_2", "rot_3"), tuple(f"f_rest_{i}" for i in range(n_rest3))):
        for a, b in zip(group, group[1:]):
            if idx[b] != idx[a] + 1:
                raise ValueError(f"{path}: properties {a},{b} are not consecutive")
    layout = dict(stride=stride, off_xyz=idx["x"], off_fdc=idx["f_dc_0"],
                  off_frest=idx.get("f_rest_0", 0), n_rest=n_rest3 // 3, off_opacity=idx["opacity"],
                  off_scale=idx["scale_0"], off_rot=idx["rot_0"])
    return raw.reshape(P, stride), layout


def activate_on_device(vertices: np.ndarray, layout: dict, device) -> Scene:
    """H2D copy of the vertex block + the fused activation/repack kernel -> Scene on `device`."""
    L = _cabi.lib()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _cabi.B200GSError("b200gs needs a CUDA device; there is no CPU fallback")
    P, stride = vertices.shape
    v = torch.from_numpy(np.ascontiguousarray(vertices, np.float32)).to(dev, non_blocking=False)
    M = 1 + layout["n_rest"]
    e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    means, shs, opac, scales, rots = e(P, 3), e(P, M, 3), e(P, 1), e(P, 3), e(P, 4)
    lay = _cabi.B200GSPlyLayout(**{k: int(layout[k]) for k, _ in _cabi.B200GSPlyLayout._fields_})
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        _cabi.check(L.b200gs_ply_activate(C.c_int32(P), p(v), C.byref(lay), p(means), p(shs), p(opac), p(scales),
                                          p(rots), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    deg = {1: 0, 4: 1, 9: 2, 16: 3}.get(M)
    if deg is None:
        raise ValueError(f"unsupported SH coefficient count {M}")
    return Scene(means, shs, opac, scales, rots, deg)


def load_gaussian_ply(path: str, device="cuda") -> Scene:
    vertices, layout = read_gaussian_ply(path)
    return activate_on_device(vertices, layout, device)
