"""Importable alias of the `strange-attractor-renderer_b200/` package (a hyphen cannot be
imported).  All code lives there; this file only redirects the package path."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          "strange-attractor-renderer_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
