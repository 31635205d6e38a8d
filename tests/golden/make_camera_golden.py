"""Generates tests/golden/camera_golden.json by IMPORTING the Python reference in this container
(/root/reference, read-only) -- it cannot travel to the GPU box, so the vectors are committed.

Pinned by these vectors:
  * the OpenGL camera-to-world + intrinsics convention of the reference's six segmenter views
    (Articulation/openbox_output/segmentation/camera_params.json, produced by
    Articulation/segmentation/interactive_segmenter.py:262-313),
  * its vertex projection (interactive_segmenter.py:1436-1460, `_project_vertices_to_2d`),
  * the Nerfstudio-format projection helper (Articulation/utils/nerf2physic_utils.py:10-23),
  * the hinge axis/origin of the articulated sample object (openbox_output/urdf/metadata.json:8-30).

Run:  python tests/golden/make_camera_golden.py
"""
import json
import os
import sys
import types
from unittest import mock

import numpy as np

REF = "/root/reference/Articulation"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "camera_golden.json")


def main():
    # the segmenter module imports GUI / model packages that are absent here; its projection
    # method uses none of them, so stub the imports and call the method unbound.
    for name in ("trimesh", "open3d", "cv2", "clip", "openai", "segment_anything",
                 "segment_anything.utils", "segment_anything.utils.transforms", "matplotlib",
                 "matplotlib.pyplot", "torch_scatter", "dotenv", "autoseg", "autoseg.utils",
                 "autoseg.utils.point_utils", "autoseg.utils.mesh_utils"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=name)
    sys.path.insert(0, REF)
    from utils import nerf2physic_utils as n2p
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_interactive_segmenter",
                                                  os.path.join(REF, "segmentation", "interactive_segmenter.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    project = mod.InteractiveSegmenter._project_vertices_to_2d

    cams = json.load(open(os.path.join(REF, "openbox_output", "segmentation", "camera_params.json")))
    meta = json.load(open(os.path.join(REF, "openbox_output", "urdf", "metadata.json")))
    rng = np.random.default_rng(20261017)
    out = {"views": {}, "hinge": meta["hinge"], "joint_limits": meta["joint_limits"]}
    for name, cp in cams.items():
        K = np.asarray(cp["intrinsics"], np.float64)
        c2w = np.asarray(cp["c2w"], np.float64)
        # points in front of the camera: a box around the look-at point (camera looks down -Z)
        fwd = -c2w[:3, 2]
        centre = c2w[:3, 3] + fwd * np.linalg.norm(c2w[:3, 3]) * 0.9
        pts = centre[None] + rng.uniform(-0.8, 0.8, size=(24, 3))
        uv = project(None, pts, K, c2w)
        uv_ns = n2p.project_3d_to_2d(pts, np.linalg.inv(c2w), K)
        out["views"][name] = {
            "intrinsics": K.tolist(), "c2w": c2w.tolist(), "resolution": cp["resolution"],
            "points": pts.tolist(), "uv_segmenter": np.asarray(uv).tolist(),
            "uv_nerfstudio": np.asarray(uv_ns).tolist(),
        }
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
