"""Generates the Nerfstudio wire-format fixtures and their known answers by IMPORTING the Python reference in this
container (/root/reference, read-only; it cannot travel to the GPU box, so inputs and outputs are committed):

  tests/golden/ns_transforms_global.json      synthetic transforms.json, one global intrinsics block
  tests/golden/ns_transforms_perframe.json    the same cameras with per-frame fl_x/fl_y/cx/cy/w/h
  tests/golden/ns_dataparser_transforms.json  synthetic dataparser_transforms.json (3x4 transform + scale)
  tests/golden/nerfstudio_golden.json         what the reference reads out of them and where it projects points

Reference code exercised (Articulation/utils/nerf2physic_utils.py): parse_transforms_json :26-52 (both
`different_Ks` modes, with and without `return_w2c`), parse_dataparser_transforms_json :55-61,
project_3d_to_2d :10-23, and the dataparser convention of load_ns_point_cloud :64-73 (a Nerfstudio-space point is
brought back to the transforms.json frame with inv([transform; 0 0 0 1/scale]) and the homogeneous divide that
open3d's PointCloud.transform applies -- open3d itself is not installed, that one line is restated with numpy).

Run:  python tests/golden/make_nerfstudio_golden.py
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference/Articulation"
HERE = os.path.dirname(os.path.abspath(__file__))


def look_at_gl(eye, target, up):
    eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
    f = target - eye; f /= np.linalg.norm(f)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = s, u, -f, eye     # OpenGL: camera looks down -Z
    return c2w


def main():
    # import the one module by path: the package's __init__ pulls in trimesh / open3d, which are absent here
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_nerf2physic_utils", os.path.join(REF, "utils", "nerf2physic_utils.py"))
    n2p = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(n2p)
    rng = np.random.default_rng(20261018)
    W, H = 640, 480
    frames_g, frames_p = [], []
    for i in range(5):
        eye = np.array([2.5 * np.cos(0.9 * i), 2.5 * np.sin(0.9 * i), 0.8 + 0.2 * i])
        c2w = look_at_gl(eye, (0.1, -0.05, 0.2), (0, 0, 1))
        fl = 520.0 + 15.0 * i
        frames_g.append({"file_path": f"images/frame_{i:05d}.png", "transform_matrix": c2w.tolist()})
        frames_p.append({"file_path": f"images/frame_{i:05d}.png", "transform_matrix": c2w.tolist(),
                         "fl_x": fl, "fl_y": fl * 1.01, "cx": W / 2 + 3.5 - i, "cy": H / 2 - 6.25 + 2 * i, "w": W, "h": H})
    tg = {"camera_model": "OPENCV", "fl_x": 540.0, "fl_y": 545.0, "cx": W / 2 + 4.5, "cy": H / 2 - 7.75, "w": W, "h": H,
          "frames": frames_g}
    tp = {"camera_model": "OPENCV", "frames": frames_p}
    a = 0.35
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]]) @ \
        np.array([[1, 0, 0], [0, np.cos(0.2), -np.sin(0.2)], [0, np.sin(0.2), np.cos(0.2)]])
    dp = {"transform": np.concatenate([R, np.array([[0.3], [-0.2], [0.1]])], 1).tolist(), "scale": 0.37}
    paths = {"global": os.path.join(HERE, "ns_transforms_global.json"), "perframe": os.path.join(HERE, "ns_transforms_perframe.json"),
             "dataparser": os.path.join(HERE, "ns_dataparser_transforms.json")}
    json.dump(tg, open(paths["global"], "w"), indent=1)
    json.dump(tp, open(paths["perframe"], "w"), indent=1)
    json.dump(dp, open(paths["dataparser"], "w"), indent=1)

    out = {}
    c2ws, K = n2p.parse_transforms_json(paths["global"])
    w2cs, K2 = n2p.parse_transforms_json(paths["global"], return_w2c=True)
    c2ws_p, Ks = n2p.parse_transforms_json(paths["perframe"], different_Ks=True)
    ns_transform, scale = n2p.parse_dataparser_transforms_json(paths["dataparser"])
    M = np.concatenate([ns_transform, np.array([[0, 0, 0, 1 / scale]])], 0)        # load_ns_point_cloud :69-70
    inv_M = np.linalg.inv(M)
    pts_ns = rng.uniform(-0.25, 0.25, size=(32, 3)) + np.array([0.1, 0.0, 0.1])
    hom = np.concatenate([pts_ns, np.ones((32, 1))], 1) @ inv_M.T
    pts_orig = hom[:, :3] / hom[:, 3:4]                                             # open3d transform: divide by w
    out["global"] = {"K": np.asarray(K).tolist(), "c2w": [np.asarray(c).tolist() for c in c2ws],
                     "w2c": [np.asarray(w).tolist() for w in w2cs],
                     "uv": [n2p.project_3d_to_2d(pts_orig, w, K).tolist() for w in w2cs]}
    out["perframe"] = {"K": [np.asarray(k).tolist() for k in Ks],
                       "uv": [n2p.project_3d_to_2d(pts_orig, np.linalg.inv(c), k).tolist() for c, k in zip(c2ws_p, Ks)]}
    out["dataparser"] = {"transform": np.asarray(ns_transform).tolist(), "scale": float(scale)}
    out["points_ns"] = pts_ns.tolist()
    out["points_orig"] = pts_orig.tolist()
    out["image"] = [W, H]
    json.dump(out, open(os.path.join(HERE, "nerfstudio_golden.json"), "w"), indent=1)
    print("wrote", sorted(paths.values()), "and nerfstudio_golden.json")


if __name__ == "__main__":
    main()
