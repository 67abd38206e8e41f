"""Import alias: the package lives in the directory `mcaq-yolo_b200/` (not a valid Python
identifier), this shim makes it importable as `mcaq_yolo_b200`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mcaq-yolo_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
