"""Alias package: ``import diff_gaussian_rasterization`` resolves to the B200 rasterizer, so an
inria-style ``render()`` or a Nerfstudio-side adapter runs unchanged."""
from robosimgs_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,
                                       RasterizationSettings, rasterize_gaussians, _RasterizeGaussians)

__all__ = ["GaussianRasterizationSettings", "RasterizationSettings", "GaussianRasterizer",
           "rasterize_gaussians"]
