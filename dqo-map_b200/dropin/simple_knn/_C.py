"""Drop-in for `simple_knn._C` (`from simple_knn._C import distCUDA2`, SLAM/gaussian_pointcloud.py:7)."""
from dqo_map_b200.knn import distCUDA2  # noqa: F401
