"""Import shim: makes `import cedarsim.jl_b200` resolve to the `cedarsim.jl_b200/` directory
at the repo root (a directory name with a dot cannot be imported directly)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "cedarsim.jl_b200")
if "cedarsim.jl_b200" not in _sys.modules:
    _spec = _ilu.spec_from_file_location(
        "cedarsim.jl_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
    )
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules["cedarsim.jl_b200"] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = _sys.modules["cedarsim.jl_b200"]
