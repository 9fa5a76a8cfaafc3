"""Importable name of the `strange-attractor-renderer_b200/` package (a hyphen cannot appear in an
import statement).  All code lives in that directory; this module loads its __init__ under the
importable name and then IS that package."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "strange-attractor-renderer_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
