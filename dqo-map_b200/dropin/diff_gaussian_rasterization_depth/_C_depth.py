"""Stand-in for the pybind module `_C_depth` (RAST/ext.cpp:15-19): same three functions, same signatures."""
from dqo_map_b200.rasterizer import mark_visible, rasterize_gaussians, rasterize_gaussians_backward  # noqa: F401
