"""Importable alias for the package directory whose (mandated) name is not a Python identifier:
``enhanced-3d-reconstruction-in-colonoscopy-using-monocular-depth-and-pose-estimation_b200/``.
``import dav2_b200`` executes that directory's ``__init__.py`` with ``__path__`` pointing at it, so
``dav2_b200.dpt`` etc. resolve to the real sources (and the in-tree ``libdav2_b200.so``)."""
import os as _os

_REAL = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "enhanced-3d-reconstruction-in-colonoscopy-using-monocular-depth-and-pose-estimation_b200",
)
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
