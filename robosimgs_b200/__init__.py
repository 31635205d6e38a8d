"""robosimgs_b200 -- B200-native (sm_100a) 3D Gaussian Splatting rasterizer behind the
GaussianRasterizer / GaussianRasterizationSettings operator surface.

Only what the hot path needs lives here: csrc/ (CUDA kernels + C ABI), the ctypes binding, the
PyTorch operator mirror, camera/scene helpers for the BASELINE configs and the camera-sharded sweep.
"""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, RasterizationSettings,
                         PairCapacityExceeded, export_rgb8, rasterize_gaussians)
from ._cabi import B200GSError, LIB_PATH

__all__ = ["GaussianRasterizationSettings", "RasterizationSettings", "GaussianRasterizer",
           "rasterize_gaussians", "export_rgb8", "B200GSError", "PairCapacityExceeded", "LIB_PATH"]
__version__ = "0.1.0"
