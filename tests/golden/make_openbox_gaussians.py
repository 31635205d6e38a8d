"""Generates tests/golden/openbox_surface_samples.npz from the reference's sample object in THIS container
(/root/reference, read-only; it cannot travel to the GPU box, so the derived samples are committed):

  /root/reference/Articulation/openbox_output/urdf/body_centered.glb   (8416 vertices, 16607 faces)
  /root/reference/Articulation/openbox_output/urdf/lid_centered.glb    (4410 vertices,  8393 faces)
  /root/reference/Articulation/openbox_output/urdf/metadata.json       (hinge axis, joint limits)

33 000 + 17 000 points sampled uniformly by area with their face normals -- the object Gaussians of BASELINE config C5
 -- through the product's own loaders (robosimgs_b200.compositor.load_glb_mesh / gaussians_on_mesh /
load_urdf_metadata), so tests/test_compositor.py can also check on this box that the loaders still reproduce the
fixture from the reference files.   Run:  python tests/golden/make_openbox_gaussians.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
URDF_DIR = "/root/reference/Articulation/openbox_output/urdf"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "openbox_surface_samples.npz")
N_BODY, N_LID, SEED = 33_000, 17_000, 5


def generate():
    from robosimgs_b200 import compositor as cp
    meta = cp.load_urdf_metadata(os.path.join(URDF_DIR, "metadata.json"))
    bv, bf = cp.load_glb_mesh(os.path.join(URDF_DIR, meta["files"]["body_mesh"]))
    lv, lf = cp.load_glb_mesh(os.path.join(URDF_DIR, meta["files"]["lid_mesh"]))
    bp, bn = cp.gaussians_on_mesh(bv, bf, N_BODY, SEED)
    lp, ln_ = cp.gaussians_on_mesh(lv, lf, N_LID, SEED + 1)
    return dict(body_pts=bp.astype(np.float32), body_nrm=bn.astype(np.float16), lid_pts=lp.astype(np.float32),
                lid_nrm=ln_.astype(np.float16), axis=meta["axis"], limits=np.asarray(meta["limits"]),
                mesh_counts=np.asarray([len(bv), len(bf), len(lv), len(lf)]),
                body_bounds=np.stack([bv.min(0), bv.max(0)]), lid_bounds=np.stack([lv.min(0), lv.max(0)]))


if __name__ == "__main__":
    d = generate()
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT), "bytes", d["mesh_counts"])
