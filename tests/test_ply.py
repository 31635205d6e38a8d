"""CPU tests of the .ply host logic (header/layout parsing, pre-activation round trip)."""
import numpy as np
import pytest
import torch

from robosimgs_b200 import ply
from robosimgs_b200.scenes import cube_scene


def test_write_read_layout_and_preactivation(tmp_path):
    sc, _ = cube_scene(P=257, seed=3, degree=3)
    path = str(tmp_path / "scene.ply")
    ply.write_gaussian_ply(path, sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations)
    v, lay = ply.read_gaussian_ply(path)
    assert v.shape == (257, 62) and lay["stride"] == 62 and lay["n_rest"] == 15
    assert (lay["off_xyz"], lay["off_fdc"], lay["off_frest"], lay["off_opacity"], lay["off_scale"], lay["off_rot"]) == \
        (0, 6, 9, 54, 55, 58)
    assert np.allclose(v[:, 0:3], sc.means3D.numpy())
    assert np.allclose(1 / (1 + np.exp(-v[:, 54].astype(np.float64))), sc.opacities.numpy()[:, 0], atol=1e-6)
    assert np.allclose(np.exp(v[:, 55:58].astype(np.float64)), sc.scales.numpy(), rtol=1e-6)
    # f_rest is channel-major in the file
    rest = v[:, 9:54].reshape(257, 3, 15).transpose(0, 2, 1)
    assert np.allclose(rest, sc.shs.numpy()[:, 1:, :])


def test_degree0_layout_and_errors(tmp_path):
    sc, _ = cube_scene(P=10, seed=3, degree=0)
    path = str(tmp_path / "d0.ply")
    ply.write_gaussian_ply(path, sc.means3D, sc.shs, sc.opacities, sc.scales, sc.rotations)
    v, lay = ply.read_gaussian_ply(path)
    assert v.shape == (10, 17) and lay["n_rest"] == 0
    bad = tmp_path / "bad.ply"
    bad.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nend_header\n0\n")
    with pytest.raises(ValueError, match="binary_little_endian"):
        ply.read_gaussian_ply(str(bad))
    trunc = tmp_path / "trunc.ply"
    trunc.write_bytes(open(path, "rb").read()[:-8])
    with pytest.raises(ValueError, match="truncated"):
        ply.read_gaussian_ply(str(trunc))
    with pytest.raises(Exception, match="CUDA"):
        ply.activate_on_device(v, lay, "cpu")
