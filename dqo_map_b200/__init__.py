"""Import shim: the package sources live in the directory `dqo-map_b200/` (the layout this repository is
required to use), which is not a valid Python identifier.  `import dqo_map_b200` resolves to that directory."""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "dqo-map_b200"))
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
