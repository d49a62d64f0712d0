"""dqo_map_b200 — B200-native rasterization / mapping hot path of DQO-MAP behind the reference's operator API.

Public surface (mirrors the reference's three native operators):
  rasterizer.GaussianRasterizationSettings / GaussianRasterizer / mark_visible / rasterize_gaussians[_backward]
  knn.distCUDA2
  map_utils.accumulate_gaussian_error
  mapping.masked_l1_loss / FusedAdam / MappingStep
  quadric.quadric_init / quadric_project / quadric_refine
  sharding.assign_objects / gather_object_table / gather_gaussians
Everything computes through libdqomap_b200.so (include/dqo_b200.h); nothing falls back to PyTorch or the CPU.
"""
__version__ = "0.1.0"
