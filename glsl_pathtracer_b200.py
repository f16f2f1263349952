"""Import shim: the package directory is named `glsl-pathtracer_b200/` (not a valid Python identifier);
this module loads it under the importable name `glsl_pathtracer_b200`."""
import importlib.util as _u, os as _os, sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "glsl-pathtracer_b200")
_spec = _u.spec_from_file_location("glsl_pathtracer_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["glsl_pathtracer_b200"] = _mod
_spec.loader.exec_module(_mod)
