"""Makes `dqo_map_b200` importable when only this `dropin/` directory is on sys.path / PYTHONPATH."""
import os
import sys

_root = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
if _root not in sys.path:
    sys.path.insert(0, _root)
