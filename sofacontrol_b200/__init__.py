"""Import shim: `soft-robot-control_b200/` (the package directory this repo's layout prescribes) is not a valid
Python identifier, so this tiny package forwards `sofacontrol_b200.*` to it.  All code lives there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "soft-robot-control_b200")
__path__.append(_real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
