"""Import alias.  The package directory carries the reference's hyphenated name (``swh-trl_b200/``), which is
not a valid Python identifier; this shim makes it importable as ``swh_trl_b200``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "swh-trl_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _os, _f
