"""Drop-in for `cuda_utils._C` (`from cuda_utils._C import accumulate_gaussian_error`, mapper.py:23)."""
from dqo_map_b200.map_utils import accumulate_gaussian_error  # noqa: F401
