"""Drop-in for the reference package `diff_gaussian_rasterization_depth` (SLAM/render.py:8-13 imports
GaussianRasterizationSettings and GaussianRasterizer from it).  Put `dqo-map_b200/dropin` on PYTHONPATH."""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import _bootstrap  # noqa: F401,E402

from dqo_map_b200.rasterizer import (  # noqa: E402,F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians_fn as rasterize_gaussians,
)
from . import _C_depth  # noqa: E402,F401
